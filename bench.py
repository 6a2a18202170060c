#!/usr/bin/env python
"""Benchmark of the PyMiniWeather hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one ``evolve()`` = two fused sweep kernels (one per direction, three RK stages each;
with ``--tune fuse=0``: six stage kernels) over the whole grid.
Workload at N=1: BASELINE config 2, thermal rising bubble, nx=2048 nz=1024, fp64.  At N>1 the
same slab (2048 x 1024 per GPU) is weak-scaled: global grid 2048*N x 1024, ring halo exchange
before every x stage.  Metric: cell-updates/s = global cells * steps / time.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

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
BYTES_PER_CELL_STEP = 512.0  # 2 sweeps x (64 + 96 + 96) B: SURVEY.md section 8d / DESIGN.md
STAGES_PER_STEP = 6


def make_params(nx_local, nz, world):
    """params of pyminiweather/__main__.py:160-195 for a slab of a domain `world` slabs wide; xlen
    grows with the number of slabs so that dx = dz and dt stay those of config 2."""
    p = dict(nx=nx_local, nz=nz, xlen=2e4 * world, zlen=1e4, hs=2, s=4, ic_type="thermal",
             max_speed=500.0, cfl=1.0)
    p["dx"] = p["xlen"] / (nx_local * world)
    p["dz"] = p["zlen"] / nz
    p["dt"] = min(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
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


# ----------------------------------------------------------------------------------------------
# CPU arms (oracle = checker code, used here only as the timed CPU baseline)
# ----------------------------------------------------------------------------------------------
def cpu_case(nx, nz):
    import numpy as np
    from oracle import numpy_oracle as no
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    p = make_params(nx, nz, 1)
    f = initialize_fields(p)
    init(f, p, MeshData(p))
    hyd = [getattr(f, n).copy() for n in ("hy_dens_cell", "hy_dens_theta_cell", "hy_dens_int",
                                          "hy_dens_theta_int", "hy_pressure_int")]
    return p, no.OracleCase(nx, nz, p["dx"], p["dz"], p["dt"], f._host[0].copy(), f._host[1].copy(), *hyd)


def time_c_oracle(case, budget_s, min_steps=2):
    """Multi-threaded C restatement (OpenMP, all host threads)."""
    from oracle import c_oracle
    c = c_oracle.COracle(case)
    c.evolve(1)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    c.evolve(min_steps)
    per = (time.perf_counter() - t0) / min_steps
    n = max(min_steps, min(2000, int(budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    c.evolve(n)
    dt = time.perf_counter() - t0
    return case.nx * case.nz * n / dt, n, dt


def time_numpy_oracle(case, nsteps):
    """Single-threaded NumPy restatement (the reference's backend minus scipy's generic
    N-D correlate, which makes it ~4x faster than the reference itself)."""
    from oracle import numpy_oracle as no
    t0 = time.perf_counter()
    for _ in range(nsteps):
        no.evolve(case)
    dt = time.perf_counter() - t0
    return case.nx * case.nz * nsteps / dt, nsteps, dt


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on this box's host cores.  The
    reference is Python/NumPy and cannot travel to the GPU box, so this is the oracle port;
    all host threads (OpenMP)."""
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    from oracle import c_oracle as co
    nz_s = NZ
    p, case = cpu_case(NX_SLAB, nz_s)
    c = co.COracle(case)
    c.evolve(1)
    t0 = time.perf_counter()
    c.evolve(2)
    per = (time.perf_counter() - t0) / 2
    # K steps must finish within a few minutes: if the full grid is too slow for the requested K,
    # each step becomes one evolve() over a horizontal band of the same workload (fewer rows)
    while per * (args.steps + args.warmup) * nz_s / NZ > 150.0 and nz_s > 64:
        nz_s //= 2
    if nz_s != NZ:
        p, case = cpu_case(NX_SLAB, nz_s)
        c = co.COracle(case)
    c.evolve(max(1, min(args.warmup, 3)))
    steps = args.steps
    t0 = time.perf_counter()
    c.evolve(steps)
    dt = time.perf_counter() - t0
    value = NX_SLAB * nz_s * steps / dt
    _, ncase = cpu_case(NX_SLAB, NZ)
    np_value, np_n, np_dt = time_numpy_oracle(ncase, 2)
    cores = host_threads()
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"thermal rising bubble nx={NX_SLAB} nz={NZ} fp64 (BASELINE config 2), CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} evolve() steps on a {NX_SLAB}x{nz_s} grid "
                                   f"({'the full workload' if nz_s == NZ else 'a band of the ' + str(NX_SLAB) + 'x' + str(NZ) + ' workload'}), "
                                   f"C/OpenMP oracle (oracle/c/pmw_oracle.c), {cores} threads",
                         "numpy_1core_value": np_value,
                         "numpy_1core_sample": f"{np_n} steps, oracle/numpy_oracle.py, 1 thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    from pyminiweather_b200 import engine
    from pyminiweather_b200._lib import PMW_BUF_STATE, PMW_BUF_TMP
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    from pyminiweather_b200.slab import SlabMesh, SlabRing

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    p = make_params(NX_SLAB, NZ, world)
    fields = initialize_fields(p)
    init(fields, p, MeshData(p) if world == 1 else SlabMesh(p, rank, world))
    hyd = [getattr(fields, n) for n in engine.HYDRO_NAMES]
    host_state = fields._host[PMW_BUF_STATE]

    solver = engine.DeviceSolver(NX_SLAB, NZ, p["dx"], p["dz"], p["dt"], device=local_rank,
                                 variant=args.variant, pow_mode=args.pow_mode, periodic_x=(world == 1))
    stream = torch.cuda.current_stream()
    solver.set_stream(stream.cuda_stream)
    solver.set_hydrostatic(*hyd)
    for kv in args.tune or []:
        k, v = kv.split("=")
        solver.set_tuning(**{k: int(v)})
    solver.upload(PMW_BUF_STATE, host_state)
    solver.upload(PMW_BUF_TMP, host_state)
    ring = None
    if world > 1:
        ring = SlabRing(solver, rank, world, lambda n: torch.zeros(n, dtype=torch.float64, device="cuda"), dist,
                        args.halo)

    def step(n):
        if ring is None:
            solver.evolve(n)
        else:
            ring.evolve(n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m0, e0 = (ring.stats() if ring else solver.stats(PMW_BUF_STATE))

    # ---- timed region A: `value` ----------------------------------------------------------
    step(args.warmup)
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = solver.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    step(args.steps)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = solver.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    cells = NX_SLAB * NZ * world
    value = cells * args.steps / (ms * 1e-3)

    # ---- timed region B: the same K steps with a CUDA-event pair around every stage kernel ---
    solver.stage_timing(True)
    barrier()
    step(args.steps)
    launch_ms, n_timed = solver.stage_timing_read()
    solver.stage_timing(False)
    barrier()

    m1, e1 = (ring.stats() if ring else solver.stats(PMW_BUF_STATE))
    finite = bool(np.isfinite(m1) and np.isfinite(e1))

    # ---- e2e: the drop-in operator call on HOST arrays (upload + evolve + download per step) ----
    e2e = None
    if rank == 0 and not args.no_e2e:
        import types
        from pyminiweather_b200 import engine as eng
        from pyminiweather_b200.solve import evolve
        eng.DEFAULTS.update(variant=args.variant, pow_mode=args.pow_mode, device=local_rank)
        p1 = make_params(NX_SLAB, NZ, 1)
        pinned = torch.empty((4, NZ + 4, NX_SLAB + 4), dtype=torch.float64, pin_memory=True)
        f1 = initialize_fields(p1)
        init(f1, p1, MeshData(p1))
        pinned.numpy()[:] = f1._host[PMW_BUF_STATE]
        foreign = types.SimpleNamespace(state=pinned.numpy(), state_tmp=None, nvariables=4)
        for n_, a_ in zip(eng.HYDRO_NAMES, [getattr(f1, n) for n in eng.HYDRO_NAMES]):
            setattr(foreign, n_, a_)
        k_e2e = max(5, min(args.steps, 50))
        for _ in range(3):
            evolve(p1, foreign, None, dt=p1["dt"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            evolve(p1, foreign, None, dt=p1["dt"])
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        nbytes = pinned.numel() * 8
        e2e = {"value": NX_SLAB * NZ * k_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": k_e2e, "ms_per_step": 1e3 * dt_e2e / k_e2e,
               "api": "pyminiweather_b200.solve.evolve(params, fields, mesh, dt) on host NumPy arrays "
                      "(pinned); single GPU"}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        _, case = cpu_case(NX_SLAB, NZ)
        c_val, c_n, c_dt = time_c_oracle(case, budget_s=args.cpu_budget)
        cpu = {"value": c_val, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"{c_n} evolve() steps of the same {NX_SLAB}x{NZ} thermal workload in {c_dt:.1f} s, "
                         f"C/OpenMP oracle on {host_threads()} host threads"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    import json as _json
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak = float(_json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = _json.load(open(tpath)).get("dram_bytes_per_launch_mean")
        except Exception:
            pass
    # Launches per step: two sweep kernels (fused path) or six stage kernels.  Algorithmic bytes are
    # SURVEY.md section 8d's stage-by-stage figure either way -- 64 + 96 + 96 = 256 B per cell per
    # directional sweep, 512 B per cell-step -- so a fused sweep, which keeps T1/T2 on chip and moves
    # only 64 B/cell, can exceed 1.0 of that roofline; `traffic` and `fused_*` show what it really
    # moves and what bounds it (the FP64 pipe).  Average launch duration = CUDA-event time of the
    # timed region / launches in it (per rank; the kernels a step launches are exactly these).
    fused = bool(solver.get_tuning("fuse")) and args.variant == "tma"
    launches_per_step = 2 if fused else STAGES_PER_STEP
    bytes_per_launch = NX_SLAB * NZ * BYTES_PER_CELL_STEP / launches_per_step
    launch_ms_region = ms / (args.steps * launches_per_step)
    achieved = bytes_per_launch / (launch_ms_region * 1e-3) / 1e9
    fused_bytes_per_launch = NX_SLAB * NZ * 64.0  # state read once + written once per sweep
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"thermal rising bubble, nx={NX_SLAB * world} nz={NZ} fp64"
                               + ((" (BASELINE config 2)" if (NX_SLAB, NZ) == (2048, 1024) else "") if world == 1 else
                                  f" = {world} x-slabs of {NX_SLAB}x{NZ}, ring halo exchange per x stage ({args.halo})"),
                   "nx": NX_SLAB * world, "nz": NZ, "variant": args.variant, "pow_mode": args.pow_mode,
                   "tiles": {k: solver.get_tuning(k) for k in ("fuse", "sweep_xp", "sweep_zt", "sweep_lz", "x_tr", "x_p", "z_cfg")},
                   "l2": f"no flush: working set = 3 state buffers x {4 * (NZ + 4) * (NX_SLAB + 4) * 8 / 1e6:.1f} MB "
                         "per GPU > 126 MB L2 (inputs larger than L2)",
                   "state_finite_after_run": finite,
                   "mass_rel_change": (m1 - m0) / m0, "energy_rel_change": (e1 - e0) / e0},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "peak_source": peak_src,
                     "kernel": ("sweep_x / sweep_z (one launch = one directional sweep = three RK stages over the slab)"
                                if fused else "stage_x_tma / stage_z_tma (one launch = one RK stage over the slab)"),
                     "algorithmic_bytes_per_launch": bytes_per_launch,
                     "algorithmic_bytes_definition": "SURVEY 8d, stage by stage: 256 B per cell per sweep (512 B per cell-step)",
                     "launch_ms_mean": launch_ms_region,
                     "launches_in_timed_region_per_rank": args.steps * launches_per_step,
                     "fused_min_bytes_per_launch": fused_bytes_per_launch if fused else None,
                     "fused_hbm_frac": (fused_bytes_per_launch / (launch_ms_region * 1e-3) / 1e9 / peak) if fused else None,
                     "note": ("fused sweeps move 64 B/cell per sweep instead of 256 and are bound by the FP64 pipe "
                              "(profiles/: sm__inst_executed_pipe_fp64), so frac is measured against the "
                              "stage-by-stage traffic the reference algorithm implies, not against what the kernel moves")
                             if fused else None,
                     # second pass over the same K steps with an event pair around every stage kernel
                     # (serialises the launches: no programmatic dependent launch overlap)
                     "event_pair_launch_ms_mean": launch_ms, "event_pair_launches": n_timed,
                     "event_pair_frac": (bytes_per_launch / (launch_ms * 1e-3) / 1e9 / peak) if launch_ms > 0 else None},
        "cpu_baseline": cpu,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="tma", choices=["tma", "direct"])
    ap.add_argument("--pow-mode", dest="pow_mode", default="background", choices=["background", "libdevice"])
    ap.add_argument("--tune", action="append", help="key=value tile tuning (x_tr, x_p, z_cfg)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="slab halo exchange at N>1: peer-memory stores from the stage kernels, or NCCL send/recv")
    ap.add_argument("--nx", type=int, default=None, help="per-GPU slab width (default 2048 = BASELINE config 2)")
    ap.add_argument("--nz", type=int, default=None, help="grid height (default 1024)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline work")
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
