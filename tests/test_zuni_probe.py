"""The universal z iteration (csrc/pmw_zuni.cuh, build macro PMW_ZSWEEP_UNIVERSAL) on the CPU: its control
flow -- window rotation, activity masks, wall rebuilds, clamped row stream -- is compiled for the host with
plain-C++ policies (tools/zuni_probe/probe.cpp) and one fused z sweep is compared with three z stages of the
NumPy oracle for every segment height, with walls, ragged last segments and the gravity-wave forcing."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_universal_z_iteration_control_flow_matches_oracle():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "zuni_probe", "run_probe.py")],
                         capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "agrees with the oracle" in res.stdout
