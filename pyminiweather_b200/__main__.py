"""Command-line driver: the reference's time loop (pyminiweather/__main__.py:183-248) on the
device-resident fields.  Same flags, same log lines, same dump format
(``np.savetxt(state.reshape(-1, nx+4), delimiter=",")`` and ``<name>_svars.<ext>``), so
tools/make_images.py of the reference still reads the output.  The state crosses PCIe only for
``--output-freq`` dumps; the two diagnostics are reduced on the device.

    python -m pyminiweather_b200 --nx 2048 --nz 1024 --nsteps 1000 --ic-type thermal
"""
from __future__ import annotations

import argparse
import logging
import sys
from pathlib import Path

import numpy as np

from .data import initialize_fields
from .ics import init, init_device
from .ics.initial_conditions import IC_TYPES
from .mesh import MeshData
from .post import compute_solution_variables, compute_stats
from .solve import evolve
from .utils.timing import TimedCodeBlock

# flag, type, default, help   (pyminiweather/__main__.py:28-157)
_FLAGS = [
    ("--nx", int, 200, "Number of points in x-direction"),
    ("--nz", int, 100, "Number of points in z-direction"),
    ("--xlen", float, 2e4, "Length of domain in x-direction"),
    ("--zlen", float, 1e4, "Length of domain in z-direction"),
    ("--nsteps", int, 10, "Number of time steps"),
    ("--dt", float, None, "Time step size (if none, calculated using CFL)"),
    ("--nwarmups", int, 0, "Number of warm-up time steps"),
    ("--max-speed", float, 500.0, "Assumed maximum speed used for the time step"),
    ("--cfl", float, 1.0, "CFL number used for the time step"),
    ("--output-freq", int, -1, "Write the solution every this many steps (disabled by default)"),
    ("--app-filename", str, "PyMiniWeatherData.txt", "Output file of the solution variables"),
]


def get_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="pyminiweather_b200")
    for flag, typ, default, text in _FLAGS:
        ap.add_argument(flag, type=typ, default=default, dest=flag[2:].replace("-", "_"),
                        help=f"{text} (default: {default})")
    ap.add_argument("--ic-type", type=str, default="thermal", choices=list(IC_TYPES), dest="ic_type")
    ap.add_argument("--hs", type=int, default=2, choices=[2], dest="hs")
    ap.add_argument("--s", type=int, default=4, choices=[4], dest="s")
    ap.add_argument("--device-init", action="store_true", default=False, dest="device_init",
                    help="integrate the initial condition on the GPU (init_state_kernel) instead of the "
                         "reference's host NumPy quadrature (not a reference flag)")
    ap.add_argument("--verbose", action="store_true", default=False, dest="verbose")
    ap.add_argument("--app-log-file", type=Path, default=None, metavar="FILE", dest="app_log_file")
    return ap


def get_params_from_args(args) -> dict:
    p = dict(vars(args))
    p["dx"] = p["xlen"] / p["nx"]
    p["dz"] = p["zlen"] / p["nz"]
    if p["dt"] is None:
        p["dt"] = np.minimum(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
    return p


def _logger(filename):
    log = logging.getLogger("pyminiweather")
    handler = logging.StreamHandler() if filename is None else logging.FileHandler(filename, mode="w")
    handler.setFormatter(logging.Formatter("%(message)s"))
    log.addHandler(handler)
    log.setLevel(logging.INFO)
    logging.getLogger("pyminiweather.log").addHandler(handler)
    logging.getLogger("pyminiweather.log").setLevel(logging.INFO)
    return log


def _dump(params, fields, append: bool):
    mode = "a" if append else "w"
    stem, ext = params["app_filename"].split(".")
    state = fields.state  # device -> host
    with open(params["app_filename"], mode) as fh:
        np.savetxt(fh, state.reshape(-1, state.shape[-1]), delimiter=",")
    with open(f"{stem}_svars.{ext}", mode) as fh:
        data = compute_solution_variables(params, fields)
        np.savetxt(fh, data.reshape(-1, data.shape[-1]), delimiter=",")


def main(argv=None) -> int:
    args, _ = get_parser().parse_known_args(argv)
    params = get_params_from_args(args)
    log = _logger(params["app_log_file"])
    if params["verbose"]:
        for k, v in params.items():
            log.info(f"{k:25s} {v}")

    with TimedCodeBlock(label="Elapsed time for initialization"):
        fields = initialize_fields(params)
        mesh = MeshData(params)
        (init_device if params.get("device_init") else init)(fields, params, mesh)

    mass0, energy0 = compute_stats(params, fields)
    log.info(f"Start: total_mass, total_energy: {mass0}, {energy0}")
    sync = lambda: fields.device(params).synchronize()  # noqa: E731

    if params["nwarmups"]:
        with TimedCodeBlock(label="Elapsed time for warmups", sync=sync):
            for _ in range(params["nwarmups"]):
                evolve(params, fields, mesh, dt=params["dt"])

    touched = False
    with TimedCodeBlock(label="Elapsed time for timestepping", sync=sync):
        for istep in range(params["nsteps"]):
            if params["output_freq"] > 0 and (istep + 1) % params["output_freq"] == 0:
                log.info(f"Step: {istep}, max(rho*t): {fields.state[3].max()}")
                _dump(params, fields, touched)
                touched = True
            evolve(params, fields, mesh, dt=params["dt"])

    mass1, energy1 = compute_stats(params, fields)
    log.info(f"End: total_mass, total_energy: {mass1}, {energy1}")
    log.info(f"Relative change in total_mass, total_energy: {(mass1 - mass0) / mass0}, "
             f"{(energy1 - energy0) / energy0}")
    fields.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
