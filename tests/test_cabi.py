"""The C ABI boundary without a GPU: the library loads, exports every function include/pmw.h
declares (and nothing the binding expects is missing), and fails loudly -- no CPU fallback --
when asked to compute without a CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "pmw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmw_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from pyminiweather_b200 import _lib
    decl = declared_functions()
    assert len(decl) >= 35
    assert sorted(_lib.SIGNATURES) == decl, (set(decl) ^ set(_lib.SIGNATURES))


def test_library_exports_every_declared_symbol():
    from pyminiweather_b200 import _lib
    lib = _lib.load()  # raises if the .so is missing or lacks a symbol of SIGNATURES
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pmw_[a-z0-9_]+)", out))
    assert set(declared_functions()) <= exported
    assert lib.pmw_version() >= 100
    assert isinstance(lib.pmw_last_error(), bytes)


def test_library_is_plain_c_abi_without_torch_or_libcuda_dependency():
    from pyminiweather_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcuda.so" not in out and "python" not in out.lower()


def test_sass_contains_tma_and_is_sm100a():
    """The production kernels really are TMA kernels for sm_100a (UTMALDG in SASS)."""
    from pyminiweather_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lst = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in lst
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "SYNCS" in sass  # TMA tiled loads completing on mbarriers


def test_no_cpu_fallback_without_a_device():
    """On a machine without a GPU every operator must raise, not compute on the host."""
    import numpy as np
    from pyminiweather_b200._lib import PmwError
    from pyminiweather_b200.engine import DeviceSolver
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(PmwError):
        DeviceSolver(32, 16, 1.0, 1.0, 0.1)
    from helpers import make_params
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.solve import evolve
    p = make_params(32, 16)
    f = initialize_fields(p)
    with pytest.raises(PmwError):
        evolve(p, f, None, dt=p["dt"])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyminiweather_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, fn)
