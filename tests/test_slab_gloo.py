"""World-size-2 (and 3) test of the x-slab ring on the gloo backend, CPU only.

The device solver is replaced by a NumPy stand-in built from the oracle's operators, so what is
under test is the host logic of pyminiweather_b200/slab.py: message pairing on the ring (including
the world=2 case where both neighbours are the same peer), which buffer is exchanged before which
stage, the sweep-order flag, and the all-reduced diagnostics.  The slab run must reproduce the
single-domain oracle bit for bit.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import HYDRO, new_case  # noqa: E402
from oracle import numpy_oracle as no  # noqa: E402
from pyminiweather_b200.slab import SlabMesh, SlabRing  # noqa: E402


def _view(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))


class NumpySlabSolver:
    """Host stand-in for DeviceSolver(periodic_x=False): same methods, oracle arithmetic."""

    def __init__(self, case):
        self.c = case
        self.S, self.T = case.state, case.state_tmp
        self.reverse_direction = False
        self.halo_len = 4 * case.nz * 2

    def pack_halo_x(self, buf, to_left, to_right):
        s, nx = (self.S, self.T)[buf], self.c.nx
        _view(to_left, self.halo_len)[:] = s[:, 2:-2, 2:4].reshape(-1)
        _view(to_right, self.halo_len)[:] = s[:, 2:-2, nx:nx + 2].reshape(-1)

    def unpack_halo_x(self, buf, from_left, from_right):
        s, nx, nz = (self.S, self.T)[buf], self.c.nx, self.c.nz
        s[:, 2:-2, 0:2] = _view(from_left, self.halo_len).reshape(4, nz, 2)
        s[:, 2:-2, nx + 2:nx + 4] = _view(from_right, self.halo_len).reshape(4, nz, 2)

    def evolve_stage(self, direction, rk, dt=None):
        c = self.c
        dt = c.dt if dt is None else dt
        forcing = self.S if rk == 1 else self.T
        if direction == no.DIR_X:  # halos were filled by the exchange: no set_bc_x here
            vals, d3 = no.interpolate_x(c, forcing)
            tend = no.compute_tend_x(c, no.compute_flux_x(c, vals, d3))
        else:
            no.set_bc_z(c, forcing)
            vals, d3 = no.interpolate_z(c, forcing)
            tend = no.compute_tend_z(c, no.compute_flux_z(c, vals, d3), forcing)
        out = self.S if rk == 3 else self.T
        out[:, 2:-2, 2:-2] = self.S[:, 2:-2, 2:-2] + (dt / (3, 2, 1)[rk - 1]) * tend

    def stats_device(self, buf, ptr):
        m, e = no.compute_stats(self.c, (self.S, self.T)[buf])
        _view(ptr, 2)[:] = (m, e)


def _slab_case(whole, rank, world):
    nxl = whole.nx // world
    cols = slice(rank * nxl, (rank + 1) * nxl + 4)
    return no.OracleCase(nxl, whole.nz, whole.dx, whole.dz, whole.dt, whole.state[:, :, cols].copy(),
                         whole.state_tmp[:, :, cols].copy(), *[getattr(whole, n).copy() for n in HYDRO])


def _worker(rank, world, port, nx, nz, nsteps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, whole = new_case(nx, nz, "collision")
        local = _slab_case(whole, rank, world)
        ring = SlabRing(NumpySlabSolver(local), rank, world, lambda n: torch.zeros(n, dtype=torch.float64), dist)
        ring.evolve(nsteps)
        m, e = ring.stats()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=local.state, tmp=local.state_tmp,
                 stats=np.array([m, e]), exchanges=ring.exchanges, reverse=ring.solver.reverse_direction)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_ring_reproduces_single_domain(world, tmp_path):
    nx, nz, nsteps = 48 * world, 20, 3
    mp.spawn(_worker, args=(world, _free_port(), nx, nz, nsteps, str(tmp_path)), nprocs=world, join=True)
    _, whole = new_case(nx, nz, "collision")
    for _ in range(nsteps):
        no.evolve(whole)
    want_stats = no.compute_stats(whole)
    nxl = nx // world
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["state"][:, 2:-2, 2:-2], whole.state[:, 2:-2, 2 + r * nxl: 2 + (r + 1) * nxl])
        assert np.array_equal(got["tmp"][:, 2:-2, 2:-2], whole.state_tmp[:, 2:-2, 2 + r * nxl: 2 + (r + 1) * nxl])
        assert int(got["exchanges"]) == 3 * nsteps            # one exchange per x stage
        assert bool(got["reverse"]) == whole.reverse_direction
        assert abs(got["stats"][0] - want_stats[0]) / want_stats[0] < 1e-14   # all-reduced totals
        assert abs(got["stats"][1] - want_stats[1]) / want_stats[1] < 1e-14


def test_single_rank_ring_is_set_bc_x():
    """A ring of one slab wraps onto itself: exactly the reference's periodic set_bc_x."""
    _, whole = new_case(40, 16, "thermal")
    ref = whole.copy()
    ring = SlabRing(NumpySlabSolver(whole), 0, 1, lambda n: torch.zeros(n, dtype=torch.float64), None)
    ring.evolve(2)
    for _ in range(2):
        no.evolve(ref)
    assert np.array_equal(whole.state[:, 2:-2, 2:-2], ref.state[:, 2:-2, 2:-2])
    assert ring.stats() == pytest.approx(no.compute_stats(ref), rel=1e-15)


def test_slab_mesh_matches_global_coordinates():
    from helpers import make_params
    from pyminiweather_b200.mesh import MeshData
    world, nxl, nz = 4, 16, 8
    pg = make_params(nxl * world, nz)
    xg, zg = MeshData(pg).get_mesh_int_ext()
    for r in range(world):
        pl = dict(pg, nx=nxl)
        xs, zs = SlabMesh(pl, r, world).get_mesh_int_ext()
        np.testing.assert_allclose(xs, xg[:, r * nxl: (r + 1) * nxl + 4], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(zs, zg[:, r * nxl: (r + 1) * nxl + 4])
