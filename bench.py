#!/usr/bin/env python
"""Benchmark of the PyMiniWeather hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one ``evolve()`` = two fused sweep kernels (one per direction, three RK stages each;
with ``--tune fuse=0``: six stage kernels) over the whole grid.

Headline workload (`value`): BASELINE config 2, thermal rising bubble, nx=2048 nz=1024, fp64, on one
GPU; at N>1 the same slab per GPU, weak-scaled (global grid 2048*N x 1024, x-slabs on a periodic ring,
halo columns pushed over NVLink from inside the x-sweep kernel).  Metric: cell-updates/s = global
cells * steps / time (CUDA events, max over ranks).

Every line also carries, measured in the same run:

* ``extra_configs`` -- the named multi-GPU shapes of BASELINE.json: config 4 (colliding thermals,
  2048 x 4096 per GPU, weak), config 3 (density current 8192 x 2048, STRONG: 8192/N columns per GPU) and
  the config-5 slab (synthetic random perturbation, 4096 x 8192 per GPU, weak), each with the time of
  the same slab / the whole grid on rank 0 alone, so that the record holds its own efficiency;
* ``parity_vs_single_gpu`` (N>1) -- a small sharded domain stepped 5 times by the ring against the same
  domain on rank 0 alone, bit for bit;
* ``e2e`` -- the metric through the public API on HOST arrays (every rank moves its own slab at N>1);
* ``api_loop`` (N=1) -- ``solve.evolve()`` called once per step on device-resident Fields, what the
  driver loop (``python -m pyminiweather_b200``) pays per step at config 2 and config 1;
* ``roofline`` -- the fused sweeps are bound by the FP64 pipe: achieved FP64 instruction rate against the
  DFMA rate measured on this GPU in this run (``pmw_fp64_peak``), with the HBM views beside it;
* ``cpu_baseline`` (N=1) -- the REFERENCE itself (oracle/_ref, a verbatim copy made by
  oracle/make_ref.py), 1 thread, timed on this box's host cores.

``--impl reference`` times the reference's own CPU implementation of the path (oracle/_ref; the C/OpenMP
port of the oracle when no copy is present) on a bounded band of the same global grid.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
# stdout when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the
# whole run and the result goes to a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _RESULT_OUT.write(json.dumps(obj) + "\n")
    _RESULT_OUT.flush()


METRIC = "cell-updates/sec (fp64, per RK3 step)"
UNIT = "cell-updates/s"
NX_SLAB, NZ = 2048, 1024  # BASELINE config 2 (per-GPU slab); --nx/--nz override it for the size studies
BYTES_PER_CELL_STEP = 512.0  # 2 sweeps x (64 + 96 + 96) B: SURVEY.md section 8d (stage by stage)
FUSED_BYTES_PER_CELL_STEP = 128.0  # what two fused sweeps must move: state read once + written once, each
STAGES_PER_STEP = 6
# FP64 instructions per cell and RK stage (57 per interface evaluation -- 28
# interpolation, 5 density and reciprocal, 15 pressure, 9 velocities and fluxes -- plus 8 per cell update)
FP64_PER_CELL_STAGE = 65


def make_params(nx_local, nz, world, ic_type="thermal", dx=None):
    """params of pyminiweather/__main__.py:160-195 for a slab of a domain `world` slabs wide.  Without
    `dx` the domain is 2e4 m per slab (config 2: dx = dz); with it xlen = nx_local * world * dx, i.e. the
    cell size of the named configuration is kept whatever the number of slabs."""
    xlen = 2e4 * world if dx is None else nx_local * world * dx
    p = dict(nx=nx_local, nz=nz, xlen=xlen, zlen=1e4, hs=2, s=4, ic_type=ic_type, max_speed=500.0, cfl=1.0)
    p["dx"] = p["xlen"] / (nx_local * world)
    p["dz"] = p["zlen"] / nz
    p["dt"] = min(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # pragma: no cover
            self.err = repr(exc)

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for bit, name in self._NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": max(self.power) if self.power else None}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------
# CPU arms.  oracle/ is checker code; it is executed here only as the timed CPU baseline.
# ----------------------------------------------------------------------------------------------
def time_reference(nx, nz, nsteps, warm=0, ic_type="thermal", xlen=2e4):
    """The reference itself (NumPy/SciPy backend, single thread by construction): its own init, its own
    evolve() in the loop of pyminiweather/__main__.py:210-237.  Returns (cell-updates/s, seconds)."""
    from oracle import reference_runner as rr
    run = rr.ReferenceRun(nx, nz, ic_type, xlen=xlen)
    if warm:
        run.evolve(warm)
    t0 = time.perf_counter()
    run.evolve(nsteps)
    dt = time.perf_counter() - t0
    return nx * nz * nsteps / dt, dt


def cpu_case(nx, nz, xlen=2e4):
    from oracle import numpy_oracle as no
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    p = make_params(nx, nz, 1)
    p["xlen"] = xlen
    p["dx"] = xlen / nx
    p["dt"] = min(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
    f = initialize_fields(p)
    init(f, p, MeshData(p))
    hyd = [getattr(f, n).copy() for n in ("hy_dens_cell", "hy_dens_theta_cell", "hy_dens_int",
                                          "hy_dens_theta_int", "hy_pressure_int")]
    return p, no.OracleCase(nx, nz, p["dx"], p["dz"], p["dt"], f._host[0].copy(), f._host[1].copy(), *hyd)


def time_c_oracle(case, budget_s, min_steps=2):
    """Multi-threaded C restatement (OpenMP); the team size is set explicitly and read back."""
    from oracle import c_oracle
    team = c_oracle.set_threads(host_threads())
    c = c_oracle.COracle(case)
    c.evolve(1)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    c.evolve(min_steps)
    per = (time.perf_counter() - t0) / min_steps
    n = max(min_steps, min(2000, int(budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    c.evolve(n)
    dt = time.perf_counter() - t0
    return case.nx * case.nz * n / dt, n, dt, team


def reference_available():
    try:
        from oracle import reference_runner as rr
        return rr.available()
    except Exception:
        return False


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, on a
    bounded band (fewer rows, full width) of the same global grid the repo arm runs at this N."""
    if rank != 0:
        return
    for k in ("LEGATE_MAX_DIM", "LEGATE_MAX_FIELDS"):
        os.environ.pop(k, None)
    nx_global = NX_SLAB * world
    steps, warm = args.steps, min(args.warmup, 3)
    workload = workload_name(world, args.halo)
    out = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": workload, "nx": nx_global, "nz": NZ}, "gpu_launches": 0}
    if reference_available():
        # ~1.3e5 cell-updates/s on one core: size the band for ~100 s in total
        budget_cells = 1.3e5 * 100.0 / (steps + warm)
        nz_s = NZ
        while nx_global * nz_s > budget_cells and nz_s > 16:
            nz_s //= 2
        value, dt = time_reference(nx_global, nz_s, steps, warm=warm, xlen=2e4 * world)
        from oracle import reference_runner as rr
        kind, cores = "reference", 1
        sample = (f"{steps} evolve() steps (+{warm} warm-up) of the reference's NumPy/SciPy backend "
                  f"(oracle/_ref, {rr.kind()}) on a {nx_global}x{nz_s} band of the {nx_global}x{NZ} grid "
                  f"(same dx, xlen; zlen kept, so dz differs), 1 thread: the reference's code path is single-threaded")
        extra = {}
        try:  # the multi-threaded port beside it, for scale
            _, case = cpu_case(NX_SLAB, min(NZ, 256))
            c_val, c_n, c_dt, team = time_c_oracle(case, budget_s=5.0)
            extra = {"port_openmp_value": c_val, "port_openmp_threads": team,
                     "port_openmp_sample": f"{c_n} steps of a {NX_SLAB}x{min(NZ, 256)} band, oracle/c (C/OpenMP), "
                                           f"{team} threads (set explicitly, read back from the runtime)"}
        except Exception as exc:  # pragma: no cover
            extra = {"port_openmp_error": repr(exc)}
    else:
        from oracle import c_oracle
        c_oracle.build()
        nz_s = NZ
        _, case = cpu_case(nx_global, nz_s, xlen=2e4 * world)
        team = c_oracle.set_threads(host_threads())
        c = c_oracle.COracle(case)
        c.evolve(max(1, warm))
        t0 = time.perf_counter()
        c.evolve(steps)
        dt = time.perf_counter() - t0
        value = nx_global * nz_s * steps / dt
        kind, cores = "port", team
        sample = (f"{steps} evolve() steps of the {nx_global}x{nz_s} grid, oracle/c/pmw_oracle.c (C/OpenMP port: "
                  f"oracle/_ref is absent), {team} threads (set explicitly, read back from the runtime)")
        extra = {}
    out.update(value=value, ms_per_step=1e3 * dt / steps)
    out["cpu_baseline"] = dict({"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                "host_threads_available": host_threads(), "cpu": cpu_model()}, **extra)
    out["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    emit(out)


def workload_name(world, halo):
    name = f"thermal rising bubble, nx={NX_SLAB * world} nz={NZ} fp64"
    if world == 1:
        return name + (" (BASELINE config 2)" if (NX_SLAB, NZ) == (2048, 1024) else "")
    return name + f" = {world} x-slabs of {NX_SLAB}x{NZ}, ring halo exchange per x sweep ({halo})"


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Bench:
    """One process = one GPU: helpers shared by the headline run and the extra configurations."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        self.torch = torch
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t)
        return float(t.item())

    # -- building a slab ----------------------------------------------------------------------
    def make_solver(self, nxl, nz, ic_type, world, rank, dx=None, synthetic_seed=None, host_init=False):
        """DeviceSolver for slab `rank` of `world` (world == 1: the whole periodic domain) with the named
        initial condition built on the device (init_state_kernel); `synthetic_seed`: config-5 recipe
        (thermal background, A_v * U(-1,1) perturbation) drawn on the host."""
        import numpy as np
        from pyminiweather_b200 import engine
        from pyminiweather_b200._lib import PMW_BUF_STATE, PMW_BUF_TMP
        from pyminiweather_b200.ics.initial import _init_profiles
        from pyminiweather_b200.ics.initial_conditions import device_spec
        from pyminiweather_b200.mesh import MeshData
        from pyminiweather_b200.slab import SlabMesh
        args = self.args
        p = make_params(nxl, nz, world, ic_type, dx)
        mesh = MeshData(p) if world == 1 else SlabMesh(p, rank, world)
        prof = types.SimpleNamespace(**{n: np.zeros(nz + (4 if "cell" in n else 1)) for n in engine.HYDRO_NAMES})
        _init_profiles(prof, ic_type, mesh)
        solver = engine.DeviceSolver(nxl, nz, p["dx"], p["dz"], p["dt"], device=self.local_rank,
                                     variant=args.variant, pow_mode=args.pow_mode, periodic_x=(world == 1))
        solver.set_stream(self.stream.cuda_stream)
        solver.set_hydrostatic(*[getattr(prof, n) for n in engine.HYDRO_NAMES])
        solver.profiles = prof  # the 1-D hydrostatic profiles this slab was built with
        for kv in args.tune or []:
            k, v = kv.split("=")
            solver.set_tuning(**{k: int(v)})
        if synthetic_seed is not None:
            rng = np.random.default_rng(synthetic_seed)
            st = np.zeros((4, nz + 4, nxl + 4))
            for v, amp in enumerate((1e-3, 1e-1, 1e-1, 1e-1)):
                st[v, 2:-2, 2:-2] = amp * rng.uniform(-1.0, 1.0, size=(nz, nxl))
            solver.upload(PMW_BUF_STATE, st)
            solver.upload(PMW_BUF_TMP, st)
        elif host_init:
            from pyminiweather_b200.data import initialize_fields
            from pyminiweather_b200.ics import init
            f = initialize_fields(p)
            init(f, p, mesh)
            solver.upload(PMW_BUF_STATE, f._host[PMW_BUF_STATE])
            solver.upload(PMW_BUF_TMP, f._host[PMW_BUF_STATE])
        else:
            bubbles, wind, bv0 = device_spec(ic_type, p["xlen"])
            x_axis, z_axis = mesh.get_axes_int_ext() if world > 1 else _axes(p)
            solver.init_state(bubbles, wind, bv0, x_axis, z_axis)
        return p, solver

    def make_ring(self, solver, world, rank):
        from pyminiweather_b200.slab import SlabRing
        if world == 1:
            return None
        torch = self.torch
        return SlabRing(solver, rank, world, lambda n: torch.zeros(n, dtype=torch.float64, device="cuda"), self.dist,
                        self.args.halo)

    # -- timing ---------------------------------------------------------------------------------
    def time_steps(self, solver, ring, steps, warmup, collective=True, clocks=False):
        """W warm-up steps, then K steps between CUDA events on the launching stream, barrier + synchronize on
        both sides; returns (ms max over ranks, launches summed over ranks, clocks dict or None)."""
        torch = self.torch
        step = (lambda n: ring.evolve(n)) if ring is not None else (lambda n: solver.evolve(n))
        sync = self.barrier if collective else torch.cuda.synchronize
        step(warmup)
        sync()
        sampler = None
        if clocks:
            sampler = ClockSampler(physical_gpu_index(self.local_rank))
            sampler.start()
        l0 = solver.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        ev0.record(self.stream)
        step(steps)
        ev1.record(self.stream)
        sync()
        ck = sampler.stop() if sampler else None
        if ck is not None and collective and self.world > 1:  # every rank's clocks: a ring runs at the pace of its slowest GPU
            allck = [None] * self.world
            self.dist.all_gather_object(allck, {"sm_mhz": ck["sm_mhz"], "reasons": ck["reasons"], "power_w_max": ck["power_w_max"]})
            ck["by_rank"] = allck
        ms = ev0.elapsed_time(ev1)
        launches = solver.launch_count - l0
        self.last_ms_by_rank = [ms]
        if collective:
            if self.world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                allt = [torch.empty_like(t) for _ in range(self.world)]
                self.dist.all_gather(allt, t)
                self.last_ms_by_rank = [float(x.item()) for x in allt]
            ms, launches = self.max_over_ranks(ms), int(self.sum_over_ranks(launches))
        if ring is not None:
            ring.check()  # raises if a halo wait ever timed out
        return ms, launches, ck


def _axes(p):
    import numpy as np
    hs, nx, nz, dx, dz = p["hs"], p["nx"], p["nz"], p["dx"], p["dz"]
    x = np.linspace(-hs * dx, (nx + hs) * dx, nx + 2 * hs, endpoint=False)
    z = np.linspace(-hs * dz, (nz + hs) * dz, nz + 2 * hs, endpoint=False)
    return x, z


EXTRA = [  # the named multi-GPU shapes of BASELINE.json
    dict(name="config 4: colliding thermals, nx=2048*N nz=4096 (16384x4096 at N=8), 2048x4096 per GPU", key="cfg4_weak",
         ic="collision", nxl=lambda w: 2048, nz=4096, dx=2e4 / 16384, scaling="weak", steps=100),
    dict(name="config 3: density current, nx=8192 nz=2048, 8192/N columns per GPU", key="cfg3_strong",
         ic="density-current", nxl=lambda w: 8192 // w, nz=2048, dx=2e4 / 8192, scaling="strong", steps=200),
    dict(name="config 5 slab: synthetic random perturbation, nx=4096*N nz=8192 (32768x8192 at N=8), 4096x8192 per GPU",
         key="cfg5_weak", ic="thermal", nxl=lambda w: 4096, nz=8192, dx=2e4 / 32768, scaling="weak", steps=60, synthetic=True),
]


def run_extra_configs(b, fp64_peak):
    """Each named shape: K steps on the ring of `world` GPUs and -- for the efficiency -- the reference size
    on rank 0 alone (the same slab for weak scaling, the whole grid for strong scaling)."""
    out = []
    world, rank = b.world, b.rank
    for cfg in EXTRA:
        rec = {"workload": cfg["name"], "scaling": cfg["scaling"], "ic_type": cfg["ic"],
               "data": "synthetic recipe of SURVEY.md 8d (default_rng(20260101 + rank), per-rank draw)"
                       if cfg.get("synthetic") else "device-side init (init_state_kernel)"}
        try:
            nxl, nz, steps = cfg["nxl"](world), cfg["nz"], cfg["steps"]
            seed = (20260101 + rank) if cfg.get("synthetic") else None
            p, solver = b.make_solver(nxl, nz, cfg["ic"], world, rank, dx=cfg["dx"], synthetic_seed=seed)
            ring = b.make_ring(solver, world, rank)
            # two timed repetitions of the K steps: the shorter one is the figure (a one-off host stall on any rank of
            # the ring shows up in a 60-150 ms region; both are kept in `ms_per_step_runs`)
            runs, by_rank = [], []
            for _ in range(2):
                runs.append(b.time_steps(solver, ring, steps, 10, clocks=True))
                by_rank.append([t / steps for t in b.last_ms_by_rank])
            ms, launches, ck = min(runs, key=lambda r: r[0])
            finite = _finite_stats(ring.stats() if ring else solver.stats())
            solver.close()
            cells = nxl * world * nz
            rec.update(nx=nxl * world, nz=nz, nx_per_gpu=nxl, steps=steps, ms_per_step=ms / steps,
                       ms_per_step_runs=[r[0] / steps for r in runs], ms_per_step_by_rank=by_rank,
                       value=cells * steps / (ms * 1e-3), unit=UNIT, gpu_launches=launches, clocks=ck,
                       state_finite_after_run=finite,
                       fp64_frac=_fp64_rate(cells, steps, ms) / fp64_peak / world if fp64_peak else None)
            if world > 1:  # the single-GPU time the efficiency refers to, rank 0 alone
                ms1 = None
                if rank == 0:
                    nx1 = nxl if cfg["scaling"] == "weak" else nxl * world
                    p1, s1 = b.make_solver(nx1, nz, cfg["ic"], 1, 0, dx=cfg["dx"], synthetic_seed=seed)
                    ms1, _, _ = b.time_steps(s1, None, steps, 10, collective=False)
                    s1.close()
                b.barrier()
                if rank == 0:
                    t1, tn = ms1 / steps, ms / steps
                    rec.update(single_gpu_ms_per_step=t1,
                               single_gpu_grid=f"{nx1}x{nz} on rank 0 alone, same run",
                               efficiency=(t1 / tn) if cfg["scaling"] == "weak" else t1 / (world * tn))
        except Exception as exc:  # pragma: no cover - a failed extra must not lose the headline
            rec["error"] = repr(exc)
        out.append(rec)
    return out


def _finite_stats(me):
    import numpy as np
    return bool(np.isfinite(me[0]) and np.isfinite(me[1]))


def _fp64_rate(cells, steps, ms):
    """warp-level FP64 instructions per second the stage arithmetic needs at this speed"""
    return FP64_PER_CELL_STAGE * STAGES_PER_STEP * cells * steps / 32.0 / (ms * 1e-3)


def parity_vs_single_gpu(b):
    """A 256*N x 128 thermal domain: 5 steps on the ring of N slabs against the same domain on rank 0
    alone (same initial state, broadcast from rank 0), every interior cell bit for bit."""
    import numpy as np
    torch = b.torch
    from pyminiweather_b200 import engine
    from pyminiweather_b200._lib import PMW_BUF_STATE, PMW_BUF_TMP
    world, rank = b.world, b.rank
    nxl, nz, nsteps = 256, 128, 5
    p = make_params(nxl * world, nz, 1)
    full = torch.zeros((4, nz + 4, nxl * world + 4), dtype=torch.float64, device="cuda")
    hyd = None
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    f = initialize_fields(p)  # the 1-D profiles are needed everywhere; the 2-D state comes from rank 0
    init(f, p, MeshData(p))
    hyd = [getattr(f, n) for n in engine.HYDRO_NAMES]
    if rank == 0:
        full.copy_(torch.from_numpy(f._host[PMW_BUF_STATE]))
    b.dist.broadcast(full, 0)
    host_full = full.cpu().numpy()
    slab = np.ascontiguousarray(host_full[:, :, rank * nxl:rank * nxl + nxl + 4])
    s = engine.DeviceSolver(nxl, nz, p["dx"], p["dz"], p["dt"], device=b.local_rank, variant=b.args.variant,
                            pow_mode=b.args.pow_mode, periodic_x=False)
    s.set_stream(b.stream.cuda_stream)
    s.set_hydrostatic(*hyd)
    s.upload(PMW_BUF_STATE, slab)
    s.upload(PMW_BUF_TMP, slab)
    ring = b.make_ring(s, world, rank)
    ring.evolve(nsteps)
    ring.check()
    mine = torch.from_numpy(s.download(PMW_BUF_STATE)[:, 2:-2, 2:-2].copy()).cuda()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    b.dist.all_gather(gathered, mine)
    s.close()
    ok = None
    if rank == 0:
        one = engine.DeviceSolver(nxl * world, nz, p["dx"], p["dz"], p["dt"], device=b.local_rank,
                                  variant=b.args.variant, pow_mode=b.args.pow_mode, periodic_x=True)
        one.set_stream(b.stream.cuda_stream)
        one.set_hydrostatic(*hyd)
        one.upload(PMW_BUF_STATE, host_full)
        one.upload(PMW_BUF_TMP, host_full)
        one.evolve(nsteps)
        want = one.download(PMW_BUF_STATE)[:, 2:-2, 2:-2]
        one.close()
        got = np.concatenate([g.cpu().numpy() for g in gathered], axis=2)
        ok = bool(np.array_equal(got, want))
    b.barrier()
    return ok, f"{nxl * world}x{nz} thermal, {nsteps} steps, {world} slabs vs rank 0 alone, interior bit for bit"


def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    from pyminiweather_b200 import engine
    from pyminiweather_b200._lib import PMW_BUF_STATE, PMW_BUF_TMP

    b = Bench(args, rank, local_rank, world)
    dist = b.dist

    # ---- headline: config 2 slab per GPU ---------------------------------------------------------
    p, solver = b.make_solver(NX_SLAB, NZ, "thermal", world, rank, host_init=True)
    ring = b.make_ring(solver, world, rank)
    m0, e0 = (ring.stats() if ring else solver.stats(PMW_BUF_STATE))
    ms, launches, clocks = b.time_steps(solver, ring, args.steps, args.warmup, clocks=True)
    cells = NX_SLAB * NZ * world
    value = cells * args.steps / (ms * 1e-3)

    # the same K steps with a CUDA-event pair around every sweep kernel (serialises the launches)
    solver.stage_timing(True)
    b.barrier()
    (ring.evolve if ring else solver.evolve)(args.steps)
    launch_ms, n_timed = solver.stage_timing_read()
    solver.stage_timing(False)
    b.barrier()
    m1, e1 = (ring.stats() if ring else solver.stats(PMW_BUF_STATE))
    finite = _finite_stats((m1, e1))
    fused = bool(solver.get_tuning("fuse")) and args.variant == "tma"
    tiles = {k: solver.get_tuning(k) for k in ("fuse", "sweep_xp", "sweep_z3", "sweep_zt", "sweep_lz", "x_tr", "x_p", "z_cfg")}
    fp64_peak, fp64_clk = solver.fp64_peak()  # warp DFMA/s of this GPU, measured now
    b.barrier()

    # ---- e2e: the public API on HOST arrays, every rank its own slab ------------------------------
    e2e = None
    if not args.no_e2e:
        k_e2e = max(5, min(args.steps, 50))
        nbytes = 4 * (NZ + 4) * (NX_SLAB + 4) * 8
        pinned = torch.empty((4, NZ + 4, NX_SLAB + 4), dtype=torch.float64, pin_memory=True)
        pinned.numpy()[:] = solver.download(PMW_BUF_STATE)
        if world == 1:
            from pyminiweather_b200 import engine as eng
            from pyminiweather_b200.solve import evolve
            eng.DEFAULTS.update(variant=args.variant, pow_mode=args.pow_mode, device=local_rank)
            foreign = types.SimpleNamespace(state=pinned.numpy(), state_tmp=None, nvariables=4)
            for n_ in eng.HYDRO_NAMES:
                setattr(foreign, n_, getattr(solver.profiles, n_))
            p1 = make_params(NX_SLAB, NZ, 1)
            call = lambda: evolve(p1, foreign, None, dt=p1["dt"])  # noqa: E731
            api = "pyminiweather_b200.solve.evolve(params, fields, mesh, dt) on host NumPy arrays (pinned)"
        else:
            host = pinned.numpy()

            def call():
                solver.upload(PMW_BUF_STATE, host)
                ring.barrier()  # peers may still read the halo columns the next push overwrites (slab.py)
                ring.evolve(1)
                solver.download(PMW_BUF_STATE, out=host)
            api = ("pyminiweather_b200.slab.SlabRing on host slabs: every rank uploads its slab, ring barrier, "
                   "evolve(1), downloads it -- per step")
        for _ in range(3):
            call()
        b.barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            call()
        b.barrier()
        dt_e2e = b.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": cells * k_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
               "d2h_bytes_per_step": nbytes * world, "steps": k_e2e, "ms_per_step": 1e3 * dt_e2e / k_e2e, "api": api,
               "bytes_note": "summed over ranks; every rank moves its own slab (host pinned memory)"}
        if world == 1:
            # the same call with the streamed step switched off (upload, step, download one after the other):
            # what overlapping H2D, the sweeps and D2H band by band buys (pmw_evolve_host)
            import pyminiweather_b200.solve.step as step_mod
            step_mod.HOST_BANDS = 1
            try:
                call()
                t0 = time.perf_counter()
                for _ in range(max(5, k_e2e // 2)):
                    call()
                dt_seq = (time.perf_counter() - t0) / max(5, k_e2e // 2)
            finally:
                step_mod.HOST_BANDS = 0
            e2e["sequential_ms_per_step"] = 1e3 * dt_seq
            e2e["sequential_value"] = cells / dt_seq
            e2e["note"] = ("evolve() on host arrays = pmw_evolve_host: the step streamed in row bands, H2D / sweeps / D2H "
                           "overlapped (bit-identical to the sequence); sequential_* = the same call with HOST_BANDS = 1")
    solver.close()

    # ---- N>1: bitwise check against a single GPU; everywhere: the named shapes ----------------------
    parity = parity_note = None
    if world > 1:
        parity, parity_note = parity_vs_single_gpu(b)
    extras = [] if args.no_extra else run_extra_configs(b, fp64_peak)

    # ---- N=1: API loop and the CPU baseline ---------------------------------------------------------
    api_loop = None
    if world == 1 and not args.no_api_loop:
        api_loop = measure_api_loop(b)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = measure_cpu_baseline(args)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_mean")
            traffic_src = "NOT measured in this run: " + tj.get("source", "profiles/ncu_traffic.json (ncu --set full capture)")
        except Exception:
            pass
    launches_per_step = 2 if fused else STAGES_PER_STEP
    launch_ms_region = ms / (args.steps * launches_per_step)
    per_gpu_cells = NX_SLAB * NZ
    # FP64 view (what binds the fused sweeps): warp-level FP64 instructions the stage arithmetic needs, per GPU
    fp64_achieved = _fp64_rate(per_gpu_cells, args.steps, ms)
    hbm_512 = per_gpu_cells * BYTES_PER_CELL_STEP / (ms / args.steps * 1e-3) / 1e9
    hbm_fused = per_gpu_cells * FUSED_BYTES_PER_CELL_STEP / (ms / args.steps * 1e-3) / 1e9
    if fused:
        roofline = {
            "bound": "fp64", "unit": "TFLOP/s",
            "achieved": fp64_achieved * 64 / 1e12, "peak": fp64_peak * 64 / 1e12, "frac": fp64_achieved / fp64_peak,
            "definition": "FP64 issue slots, FMA-equivalent: (65 FP64 instructions per cell and RK stage x 6 stages x cells "
                          "/ 32 lanes) per second x 64 flop, against the DFMA rate of this GPU measured in this run "
                          "(pmw_fp64_peak: 4 independent chains per thread, 16 warps per SM, all SMs)",
            "achieved_warp_instr_per_s": fp64_achieved, "peak_warp_instr_per_s": fp64_peak,
            "peak_warp_instr_per_clk_per_smsp": fp64_peak / (148 * 4 * fp64_clk * 1e6) if fp64_clk else None,
            "peak_clock_mhz": fp64_clk, "nominal_peak_warp_instr_per_clk_per_smsp": 0.5,
            "kernel": "sweep_x / sweep_z (one launch = one directional sweep = three RK stages over the slab)",
            "launch_ms_mean": launch_ms_region, "launches_in_timed_region_per_rank": args.steps * launches_per_step,
            "event_pair_launch_ms_mean": launch_ms, "event_pair_launches": n_timed,
            "traffic": traffic, "traffic_source": traffic_src,
            "hbm": {"peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                    "fused_min_bytes_per_cell_step": FUSED_BYTES_PER_CELL_STEP, "fused_achieved": hbm_fused,
                    "fused_frac": hbm_fused / hbm_peak,
                    "stage_by_stage_bytes_per_cell_step": BYTES_PER_CELL_STEP, "stage_by_stage_achieved": hbm_512,
                    "stage_by_stage_frac": hbm_512 / hbm_peak,
                    "note": "SURVEY 8d's 512 B per cell-step is the traffic of the stage-by-stage algorithm; the fused "
                            "sweeps keep T1/T2 on chip and move 128 B, so the 512 B figure can exceed the HBM peak -- it "
                            "is context, not the bound"},
        }
    else:
        bytes_per_launch = per_gpu_cells * BYTES_PER_CELL_STEP / launches_per_step
        achieved = bytes_per_launch / (launch_ms_region * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                    "kernel": "stage_x_tma / stage_z_tma (one launch = one RK stage over the slab)",
                    "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms_mean": launch_ms_region,
                    "event_pair_launch_ms_mean": launch_ms, "event_pair_launches": n_timed}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world, args.halo),
                   "nx": NX_SLAB * world, "nz": NZ, "variant": args.variant, "pow_mode": args.pow_mode, "tiles": tiles,
                   "l2": f"no flush: working set = 3 state buffers x {4 * (NZ + 4) * (NX_SLAB + 4) * 8 / 1e6:.1f} MB "
                         "per GPU > 126 MB L2 (inputs larger than L2)",
                   "state_finite_after_run": finite,
                   "mass_rel_change": (m1 - m0) / m0, "energy_rel_change": (e1 - e0) / e0},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extra_configs": extras,
    }
    if world > 1:
        out["parity_vs_single_gpu"] = parity
        out["parity_vs_single_gpu_case"] = parity_note
    if api_loop is not None:
        out["api_loop"] = api_loop
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def measure_api_loop(b):
    """What the reference's driver loop pays per step through the drop-in API: solve.evolve(params, fields,
    mesh, dt) once per step on device-resident Fields (python -m pyminiweather_b200 does exactly this,
    __main__.py:237 of the reference), host wall clock, device synchronised at both ends."""
    torch = b.torch
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init_device
    from pyminiweather_b200.mesh import MeshData
    from pyminiweather_b200.solve import evolve
    out = {"api": "pyminiweather_b200.solve.evolve(params, fields, mesh, dt), one call per step, device-resident Fields"}
    for key, nx, nz, n in (("config2_2048x1024", 2048, 1024, 500), ("config1_100x50", 100, 50, 2000)):
        p = make_params(nx, nz, 1)
        p["xlen"] = 2e4
        p["dx"] = 2e4 / nx
        p["dt"] = min(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
        f = initialize_fields(p)
        mesh = MeshData(p)
        init_device(f, p, mesh)
        for _ in range(20):
            evolve(p, f, mesh, dt=p["dt"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            evolve(p, f, mesh, dt=p["dt"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        solver = f._solver
        t1 = time.perf_counter()
        solver.evolve(n)
        solver.synchronize()
        dt_one = time.perf_counter() - t1
        out[key] = {"steps": n, "us_per_step": 1e6 * dt / n, "value": nx * nz * n / dt, "unit": UNIT,
                    "one_call_for_all_steps_us_per_step": 1e6 * dt_one / n}
        f.close()
    return out


def measure_cpu_baseline(args):
    """cpu_baseline at N=1: the reference itself on one core (its code path is single-threaded) -- BASELINE.md
    section 4: config 1 and the config-2 grid -- plus the OpenMP port of the oracle for scale."""
    out = None
    if reference_available():
        for k in ("LEGATE_MAX_DIM", "LEGATE_MAX_FIELDS"):
            os.environ.pop(k, None)
        from oracle import reference_runner as rr
        v2, t2 = time_reference(NX_SLAB, NZ, 2)
        n1 = 250
        v1, t1 = time_reference(100, 50, n1, warm=5)
        out = {"value": v2, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"2 evolve() steps of the same {NX_SLAB}x{NZ} thermal workload in {t2:.1f} s: the reference's own "
                         f"NumPy/SciPy backend (oracle/_ref, {rr.kind()}), 1 thread -- its code path is single-threaded",
               "config1_value": v1,
               "config1_sample": f"{n1} of the 1000 steps of BASELINE config 1 (thermal 100x50) in {t1:.1f} s, same code",
               "host_threads_available": host_threads(), "cpu": cpu_model(),
               "extrapolation": "configs 3-5 are not run on the CPU: at this rate one step of config 5 takes ~35 min "
                                "(BASELINE.md section 4)"}
    try:
        _, case = cpu_case(NX_SLAB, NZ)
        c_val, c_n, c_dt, team = time_c_oracle(case, budget_s=args.cpu_budget)
        port = {"port_openmp_value": c_val, "port_openmp_threads": team,
                "port_openmp_sample": f"{c_n} evolve() steps of the same {NX_SLAB}x{NZ} workload in {c_dt:.1f} s, "
                                      f"oracle/c (C/OpenMP restatement), {team} threads (set explicitly, read back)"}
    except Exception as exc:  # pragma: no cover
        port = {"port_openmp_error": repr(exc)}
    if out is None:
        out = {"value": port.get("port_openmp_value"), "unit": UNIT, "cores": port.get("port_openmp_threads"),
               "kind": "port", "sample": port.get("port_openmp_sample", "oracle/_ref absent and the port failed")}
    out.update(port)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="tma", choices=["tma", "direct"])
    ap.add_argument("--pow-mode", dest="pow_mode", default="background", choices=["background", "libdevice"])
    ap.add_argument("--tune", action="append", help="key=value tuning switch of the library (pmw_set_tuning)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="slab halo exchange at N>1: peer-memory stores from the sweep kernels, or NCCL send/recv")
    ap.add_argument("--nx", type=int, default=None, help="per-GPU slab width (default 2048 = BASELINE config 2)")
    ap.add_argument("--nz", type=int, default=None, help="grid height (default 1024)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the named multi-GPU shapes (extra_configs)")
    ap.add_argument("--no-api-loop", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=8.0, help="seconds of OpenMP-port work in cpu_baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global NX_SLAB, NZ
    if args.nx:
        NX_SLAB = args.nx
    if args.nz:
        NZ = args.nz

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})",
                  file=sys.stderr)
            sys.exit(2)
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
