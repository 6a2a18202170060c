"""``DeviceSolver``: one libpmw context = one x-slab of the domain on one B200.

Thin, typed wrapper over the C ABI (``include/pmw.h``); all numerics live in the CUDA
library.  The reference has no equivalent object -- its state lives in NumPy arrays inside
``Fields`` (pyminiweather/data/fields.py:7-55) and every operator allocates temporaries.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (PMW_BUF_STATE, PMW_BUF_TMP, PMW_DIR_X, PMW_DIR_Z, PMW_POW_BACKGROUND,
                   PMW_POW_LIBDEVICE, PMW_VARIANT_DIRECT, PMW_VARIANT_TMA, PmwParams, check)

HYDRO_NAMES = ("hy_dens_cell", "hy_dens_theta_cell", "hy_dens_int", "hy_dens_theta_int", "hy_pressure_int")

_VARIANTS = {"direct": PMW_VARIANT_DIRECT, "tma": PMW_VARIANT_TMA}
_POW = {"libdevice": PMW_POW_LIBDEVICE, "background": PMW_POW_BACKGROUND}

# process-wide defaults (tests and bench override them)
DEFAULTS = {"variant": "tma", "pow_mode": "background", "device": None}


def _as_f64(a, shape=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


class DeviceSolver:
    """Owns the device buffers of one slab and launches the fused stage kernels."""

    def __init__(self, nx: int, nz: int, dx: float, dz: float, dt: float, *, hs: int = 2,
                 device: int | None = None, variant: str | None = None, pow_mode: str | None = None,
                 periodic_x: bool = True):
        self._h = None
        lib = _lib.load()
        variant = variant or DEFAULTS["variant"]
        pow_mode = pow_mode or DEFAULTS["pow_mode"]
        if device is None:
            device = DEFAULTS["device"] if DEFAULTS["device"] is not None else 0
        if variant not in _VARIANTS:
            raise ValueError(f"variant must be one of {sorted(_VARIANTS)}")
        if pow_mode not in _POW:
            raise ValueError(f"pow_mode must be one of {sorted(_POW)}")
        self.nx, self.nz, self.hs = int(nx), int(nz), int(hs)
        self.dx, self.dz, self.dt = float(dx), float(dz), float(dt)
        self.variant, self.pow_mode, self.device = variant, pow_mode, int(device)
        self.shape = (4, self.nz + 2 * self.hs, self.nx + 2 * self.hs)
        p = PmwParams(self.nx, self.nz, self.hs, self.dx, self.dz, self.dt, self.device,
                      _VARIANTS[variant], _POW[pow_mode], 1 if periodic_x else 0)
        h = C.c_void_p()
        check(lib.pmw_create(C.byref(p), C.byref(h)))
        self._h = h
        self._lib = lib
        self._hydro = None
        self._source = None
        self._inflow_key = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pmw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        check(self._lib.pmw_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(self._lib.pmw_synchronize(self._h))

    # -- data -------------------------------------------------------------------------------
    def set_hydrostatic(self, hy_dens_cell, hy_dens_theta_cell, hy_dens_int, hy_dens_theta_int, hy_pressure_int):
        arrs = [_as_f64(hy_dens_cell, (self.nz + 4,), "hy_dens_cell"),
                _as_f64(hy_dens_theta_cell, (self.nz + 4,), "hy_dens_theta_cell"),
                _as_f64(hy_dens_int, (self.nz + 1,), "hy_dens_int"),
                _as_f64(hy_dens_theta_int, (self.nz + 1,), "hy_dens_theta_int"),
                _as_f64(hy_pressure_int, (self.nz + 1,), "hy_pressure_int")]
        dp = C.POINTER(C.c_double)
        check(self._lib.pmw_set_hydrostatic(self._h, *[a.ctypes.data_as(dp) for a in arrs]))
        self._hydro = [a.copy() for a in arrs]

    def set_source_w(self, field):
        """Gravity-wave forcing on rho*w ([nz, nx]) or None to clear it."""
        if field is None:
            check(self._lib.pmw_set_source_w(self._h, None))
            self._source = None
            return
        field = _as_f64(field, (self.nz, self.nx), "source_w")
        check(self._lib.pmw_set_source_w(self._h, C.c_void_p(field.ctypes.data)))
        self._source = field.copy()

    def set_inflow(self, row_mask, u_in: float = 50.0, theta_in: float = 298.0):
        """Injection configuration: uint8[nz] mask of the jet rows (``set_bc_x`` then takes the
        inflow branch, bcs.py:37,41-64), or None for periodic x."""
        if row_mask is None:
            check(self._lib.pmw_set_inflow(self._h, None, 0.0, 0.0))
            return
        m = np.ascontiguousarray(row_mask, dtype=np.uint8)
        if m.shape != (self.nz,):
            raise ValueError(f"row_mask: expected shape ({self.nz},), got {m.shape}")
        check(self._lib.pmw_set_inflow(self._h, C.c_void_p(m.ctypes.data), float(u_in), float(theta_in)))

    def init_state(self, bubbles, wind: float, bv0: float | None, x_axis, z_axis):
        """Device-side ``init`` of the 2-D state (initial.py:57-80): ``bubbles`` is a list of
        (amplitude, x0, z0, xrad, zrad); ``bv0`` selects the constant-Brunt-Vaisala background.
        Fills both device buffers, halo cells included."""
        spec = _lib.PmwIcSpec()
        if len(bubbles) > _lib.PMW_IC_MAX_BUBBLES:
            raise ValueError(f"at most {_lib.PMW_IC_MAX_BUBBLES} bubbles")
        spec.nbubbles = len(bubbles)
        for n, (amp, x0, z0, xrad, zrad) in enumerate(bubbles):
            spec.amp[n], spec.x0[n], spec.z0[n], spec.xrad[n], spec.zrad[n] = amp, x0, z0, xrad, zrad
        spec.wind = float(wind)
        spec.bvfreq = 0 if bv0 is None else 1
        spec.bv0 = 0.0 if bv0 is None else float(bv0)
        xa = _as_f64(x_axis, (self.nx + 2 * self.hs,), "x_axis")
        za = _as_f64(z_axis, (self.nz + 2 * self.hs,), "z_axis")
        dp = C.POINTER(C.c_double)
        check(self._lib.pmw_init_state(self._h, C.byref(spec), xa.ctypes.data_as(dp), za.ctypes.data_as(dp)))

    def hydro_matches(self, arrs) -> bool:
        return self._hydro is not None and all(np.array_equal(a, b) for a, b in zip(self._hydro, arrs))

    def upload(self, buf: int, host: np.ndarray, asynchronous: bool = False):
        host = _as_f64(host, self.shape, "state")
        fn = self._lib.pmw_upload_state_async if asynchronous else self._lib.pmw_upload_state
        check(fn(self._h, buf, C.c_void_p(host.ctypes.data)))

    def download(self, buf: int, out: np.ndarray | None = None, asynchronous: bool = False) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        if out.dtype != np.float64 or not out.flags.c_contiguous or tuple(out.shape) != self.shape:
            raise ValueError("download target must be a C-contiguous float64 array of the state shape")
        fn = self._lib.pmw_download_state_async if asynchronous else self._lib.pmw_download_state
        check(fn(self._h, buf, C.c_void_p(out.ctypes.data)))
        return out

    def upload_ptr(self, buf: int, host_ptr: int, asynchronous: bool = True):
        """Raw-pointer upload (pinned torch tensors)."""
        fn = self._lib.pmw_upload_state_async if asynchronous else self._lib.pmw_upload_state
        check(fn(self._h, buf, C.c_void_p(host_ptr)))

    def download_ptr(self, buf: int, host_ptr: int, asynchronous: bool = True):
        fn = self._lib.pmw_download_state_async if asynchronous else self._lib.pmw_download_state
        check(fn(self._h, buf, C.c_void_p(host_ptr)))

    # -- operators ----------------------------------------------------------------------------
    def bc_x(self, buf: int):
        check(self._lib.pmw_bc_x(self._h, buf))

    def bc_z(self, buf: int):
        check(self._lib.pmw_bc_z(self._h, buf))

    def stage(self, direction: int, init_buf: int, forcing_buf: int, out_buf: int, dt_stage: float):
        check(self._lib.pmw_stage(self._h, direction, init_buf, forcing_buf, out_buf, float(dt_stage)))

    def discrete_step(self, direction: int, init_buf: int, forcing_buf: int, out_buf: int, dt_stage: float):
        check(self._lib.pmw_discrete_step(self._h, direction, init_buf, forcing_buf, out_buf, float(dt_stage)))

    def evolve(self, nsteps: int = 1, dt: float | None = None):
        check(self._lib.pmw_evolve(self._h, int(nsteps), float(dt) if dt is not None else -1.0))

    def evolve_host(self, host: np.ndarray, dt: float | None = None, nbands: int = 0):
        """One ``evolve`` on a host array, in place (``pmw_evolve_host``): upload, the two fused sweeps and
        download run band by band, so that H2D and D2H overlap.  Same bits as upload + evolve(1) + download."""
        if not (isinstance(host, np.ndarray) and host.dtype == np.float64 and host.flags.c_contiguous
                and host.flags.writeable and tuple(host.shape) == self.shape):
            raise ValueError(f"state must be a writable C-contiguous float64 array of shape {self.shape}")
        check(self._lib.pmw_evolve_host(self._h, C.c_void_p(host.ctypes.data),
                                        float(dt) if dt is not None else -1.0, int(nbands)))

    def evolve_stage(self, direction: int, rk_stage: int, dt: float | None = None):
        check(self._lib.pmw_evolve_stage(self._h, direction, rk_stage, float(dt) if dt is not None else -1.0))

    @property
    def reverse_direction(self) -> bool:
        r = C.c_int()
        check(self._lib.pmw_get_reverse_direction(self._h, C.byref(r)))
        return bool(r.value)

    @reverse_direction.setter
    def reverse_direction(self, v: bool):
        check(self._lib.pmw_set_reverse_direction(self._h, 1 if v else 0))

    def stats(self, buf: int = PMW_BUF_STATE):
        out = (C.c_double * 2)()
        check(self._lib.pmw_stats(self._h, buf, out))
        return float(out[0]), float(out[1])

    def stats_device(self, buf: int, dev_ptr: int):
        check(self._lib.pmw_stats_device(self._h, buf, C.c_void_p(dev_ptr)))

    def solution_variables(self, buf: int = PMW_BUF_STATE) -> np.ndarray:
        out = np.empty((4, self.nz, self.nx), dtype=np.float64)
        check(self._lib.pmw_solution_variables(self._h, buf, C.c_void_p(out.ctypes.data)))
        return out

    # -- unfused operator shims ---------------------------------------------------------------------
    def interpolate(self, direction: int, buf: int, vals: np.ndarray, d3: np.ndarray):
        check(self._lib.pmw_interpolate(self._h, direction, buf, C.c_void_p(vals.ctypes.data),
                                        C.c_void_p(d3.ctypes.data)))

    def compute_flux(self, direction: int, vals: np.ndarray, d3: np.ndarray, flux: np.ndarray):
        check(self._lib.pmw_compute_flux(self._h, direction, C.c_void_p(vals.ctypes.data),
                                         C.c_void_p(d3.ctypes.data), C.c_void_p(flux.ctypes.data)))

    def compute_tend(self, direction: int, flux: np.ndarray, state_buf: int, tend: np.ndarray):
        check(self._lib.pmw_compute_tend(self._h, direction, C.c_void_p(flux.ctypes.data), state_buf,
                                         C.c_void_p(tend.ctypes.data)))

    # -- slab halo messages ----------------------------------------------------------------------
    @property
    def halo_len(self) -> int:
        return int(self._lib.pmw_halo_len(self._h))

    def pack_halo_x(self, buf: int, to_left_ptr: int, to_right_ptr: int):
        check(self._lib.pmw_pack_halo_x(self._h, buf, C.c_void_p(to_left_ptr), C.c_void_p(to_right_ptr)))

    def unpack_halo_x(self, buf: int, from_left_ptr: int, from_right_ptr: int):
        check(self._lib.pmw_unpack_halo_x(self._h, buf, C.c_void_p(from_left_ptr), C.c_void_p(from_right_ptr)))

    # -- peer-memory ring ----------------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        blob = C.create_string_buffer(256)
        check(self._lib.pmw_ipc_export(self._h, blob))
        return blob.raw

    def ipc_open(self, blob: bytes):
        ptrs = (C.c_void_p * 4)()
        buf = C.create_string_buffer(bytes(blob), 256)
        check(self._lib.pmw_ipc_open(self._h, buf, ptrs))
        return [p for p in ptrs]

    def local_ptrs(self):
        ptrs = (C.c_void_p * 4)()
        check(self._lib.pmw_local_ptrs(self._h, ptrs))
        return [p for p in ptrs]

    def connect_peers(self, left_ptrs, right_ptrs):
        l = (C.c_void_p * 4)(*left_ptrs)
        r = (C.c_void_p * 4)(*right_ptrs)
        check(self._lib.pmw_connect_peers(self._h, l, r))

    def peer_timed_out(self) -> bool:
        t = C.c_int()
        check(self._lib.pmw_peer_status(self._h, C.byref(t)))
        return bool(t.value)

    # -- tuning / introspection ---------------------------------------------------------------------
    def set_tuning(self, **kv):
        for k, v in kv.items():
            check(self._lib.pmw_set_tuning(self._h, k.encode(), int(v)))

    def get_tuning(self, key: str) -> int:
        v = C.c_int()
        check(self._lib.pmw_get_tuning(self._h, key.encode(), C.byref(v)))
        return v.value

    def buffer_info(self, buf: int):
        base, pitch, vs = C.c_void_p(), C.c_size_t(), C.c_size_t()
        check(self._lib.pmw_buffer_info(self._h, buf, C.byref(base), C.byref(pitch), C.byref(vs)))
        return base.value, pitch.value, vs.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.pmw_launch_count(self._h))

    def fp64_peak(self):
        """(warp-level DFMA instructions per second the GPU sustains, SM clock in MHz): the measured FP64
        roofline of the fused sweeps (``pmw_fp64_peak``)."""
        r, clk = C.c_double(), C.c_double()
        check(self._lib.pmw_fp64_peak(self._h, C.byref(r), C.byref(clk)))
        return r.value, clk.value

    def stage_timing(self, enable: bool):
        check(self._lib.pmw_stage_timing(self._h, 1 if enable else 0))

    def stage_timing_read(self):
        ms, n = C.c_double(), C.c_longlong()
        check(self._lib.pmw_stage_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
