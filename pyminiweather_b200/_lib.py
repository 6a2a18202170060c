"""ctypes binding of ``libpmw.so`` (C ABI declared in ``include/pmw.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``build_library()`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  Loading fails loudly when it is missing:
there is deliberately no CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libpmw.so")
SOURCES = ["pmw_api.cu"]
HEADERS = sorted(f for f in os.listdir(os.path.join(_PKG, "csrc")) if f.endswith(".cuh"))  # every header is a dependency

PMW_BUF_STATE, PMW_BUF_TMP = 0, 1
PMW_DIR_X, PMW_DIR_Z = 1, 2
PMW_VARIANT_DIRECT, PMW_VARIANT_TMA = 0, 1
PMW_POW_LIBDEVICE, PMW_POW_BACKGROUND = 0, 1


class PmwParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("nz", C.c_int), ("hs", C.c_int),
                ("dx", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
                ("device", C.c_int), ("variant", C.c_int), ("pow_mode", C.c_int),
                ("periodic_x", C.c_int)]


PMW_IC_MAX_BUBBLES = 4


class PmwIcSpec(C.Structure):
    """pmw_ic_spec (include/pmw.h): the configuration pmw_init_state integrates."""
    _fields_ = [("nbubbles", C.c_int),
                ("amp", C.c_double * PMW_IC_MAX_BUBBLES),
                ("x0", C.c_double * PMW_IC_MAX_BUBBLES), ("z0", C.c_double * PMW_IC_MAX_BUBBLES),
                ("xrad", C.c_double * PMW_IC_MAX_BUBBLES), ("zrad", C.c_double * PMW_IC_MAX_BUBBLES),
                ("wind", C.c_double), ("bvfreq", C.c_int), ("bv0", C.c_double)]


class PmwError(RuntimeError):
    """A libpmw call returned a non-zero status; the message is pmw_last_error()."""


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/pmw.h declares
SIGNATURES = {
    "pmw_last_error": (C.c_char_p, []),
    "pmw_version": (C.c_int, []),
    "pmw_create": (C.c_int, [C.POINTER(PmwParams), C.POINTER(_vp)]),
    "pmw_destroy": (C.c_int, [_vp]),
    "pmw_set_stream": (C.c_int, [_vp, _vp]),
    "pmw_synchronize": (C.c_int, [_vp]),
    "pmw_set_hydrostatic": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "pmw_set_source_w": (C.c_int, [_vp, _vp]),
    "pmw_set_inflow": (C.c_int, [_vp, _vp, C.c_double, C.c_double]),
    "pmw_init_state": (C.c_int, [_vp, C.POINTER(PmwIcSpec), _dp, _dp]),
    "pmw_upload_state": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_download_state": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_upload_state_async": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_download_state_async": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_bc_x": (C.c_int, [_vp, C.c_int]),
    "pmw_bc_z": (C.c_int, [_vp, C.c_int]),
    "pmw_stage": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "pmw_discrete_step": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "pmw_evolve": (C.c_int, [_vp, C.c_int, C.c_double]),
    "pmw_evolve_host": (C.c_int, [_vp, _vp, C.c_double, C.c_int]),
    "pmw_evolve_stage": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double]),
    "pmw_get_reverse_direction": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "pmw_set_reverse_direction": (C.c_int, [_vp, C.c_int]),
    "pmw_stats": (C.c_int, [_vp, C.c_int, _dp]),
    "pmw_stats_device": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_solution_variables": (C.c_int, [_vp, C.c_int, _vp]),
    "pmw_interpolate": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "pmw_compute_flux": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "pmw_compute_tend": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "pmw_halo_len": (C.c_size_t, [_vp]),
    "pmw_pack_halo_x": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pmw_unpack_halo_x": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pmw_ipc_export": (C.c_int, [_vp, _vp]),
    "pmw_ipc_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "pmw_local_ptrs": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pmw_connect_peers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "pmw_peer_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "pmw_set_tuning": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "pmw_get_tuning": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int)]),
    "pmw_buffer_info": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pmw_launch_count": (C.c_longlong, [_vp]),
    "pmw_stage_timing": (C.c_int, [_vp, C.c_int]),
    "pmw_stage_timing_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "pmw_fp64_peak": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add it to PATH)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_PKG, "csrc", f) for f in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "pmw.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libpmw.so in-tree for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH + ".tmp"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(_PKG, "csrc", s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports CC/CXX wrappers that are not meant for nvcc's host pass
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


def load():
    """Return the loaded library (cached).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("PMW_LIB", LIB_PATH)  # development: an alternative build of the same library
        if not os.path.exists(path):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the pyminiweather_b200 operators)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().pmw_last_error()
        raise PmwError(f"libpmw error {rc}: {msg.decode() if msg else '?'}")
