"""x-slab sharding of the domain across the GPUs of one box (one process per GPU).

The reference has no explicit decomposition: under ``legate --gpus N`` cuPyNumeric partitions
every array operation implicitly (README.md:7,30,245-249).  Here the grid is split into
``world`` contiguous slabs of ``nx/world`` columns.  z is not split, so z stages need no
communication; before every x stage each rank needs its neighbours' two edge columns of the
forcing state -- ``set_bc_x`` (pyminiweather/ics/bcs.py:35-39) generalised to a periodic ring.
Diagnostics are all-reduced (2 doubles).

Two exchange paths:

* ``peer`` (production): the stage kernels store their edge cell pairs straight into the
  neighbours' halo columns through IPC-mapped peer memory (NVLink P2P) and publish a per-stage
  epoch flag; the neighbours' x stages run their edge tiles last and make them wait for that
  epoch.  No pack/unpack kernels, no NCCL call, no host involvement per stage: the whole time
  loop is one ``pmw_evolve`` call per rank.  ``torch.distributed`` only carries the 256-byte IPC
  handles at start-up, barriers, and the 2-double all-reduce of the diagnostics.
* ``nccl`` (baseline / fallback): pack kernel -> ``batch_isend_irecv`` between ring neighbours on
  [4][nz][2] messages -> unpack kernel, before every x stage.  The same code runs on ``gloo`` with
  a host-side solver stand-in, which is how the exchange logic is tested without GPUs
  (tests/test_slab_gloo.py).
"""
from __future__ import annotations

import numpy as np

from ._lib import PMW_BUF_STATE, PMW_BUF_TMP, PMW_DIR_X, PMW_DIR_Z


class SlabMesh:
    """Coordinates of one slab in GLOBAL coordinates, with the MeshData getters ``init`` uses
    (pyminiweather/mesh.py:25-112)."""

    def __init__(self, params, rank: int, world: int):
        self.hs = params["hs"]
        self.dx, self.dz = params["dx"], params["dz"]
        self.nx_local, self.nz = params["nx"], params["nz"]  # params["nx"] is the slab width
        self.x0 = rank * self.nx_local * self.dx
        self._cache = {}

    def get_axes_int_ext(self):
        """1-D axes of the slab's array columns / rows (what the device-side init takes)."""
        hs, n = self.hs, self.nx_local
        # same spacing as linspace(-hs*dx, (nx+hs)*dx, nx+2hs, endpoint=False), shifted
        x = self.x0 + (np.arange(n + 2 * hs) - hs) * self.dx
        z = np.linspace(-hs * self.dz, (self.nz + hs) * self.dz, self.nz + 2 * hs, endpoint=False)
        return x, z

    def get_mesh_int_ext(self):
        if "ie" not in self._cache:
            self._cache["ie"] = np.meshgrid(*self.get_axes_int_ext())
        return self._cache["ie"]

    def get_mesh_vertical_cell_edges(self):
        return np.linspace(0.0, (self.nz + 1) * self.dz, self.nz + 1, endpoint=False)

    def get_mesh_vertical_cell_centers_int_ext(self):
        return np.linspace((-self.hs + 0.5) * self.dz, (self.nz + self.hs + 0.5) * self.dz,
                           self.nz + 2 * self.hs, endpoint=False)


class SlabRing:
    """Drives one slab solver through ``evolve`` with a halo exchange before every x stage.

    ``solver`` needs: evolve_stage, pack_halo_x, unpack_halo_x, stats_device/stats, halo_len and
    a ``reverse_direction`` property -- i.e. a ``DeviceSolver(periodic_x=False)`` or the NumPy
    stand-in used by the gloo tests.  ``make_buffer(n)`` returns a 1-D float64 torch tensor on the
    device the backend communicates from.
    """

    def __init__(self, solver, rank: int, world: int, make_buffer, dist=None, mode: str = "nccl"):
        self.solver, self.rank, self.world = solver, rank, world
        self.left, self.right = (rank - 1) % world, (rank + 1) % world
        self.stats_buf = make_buffer(2)
        self.dist = dist
        self.exchanges = 0
        self.mode = mode
        if mode == "peer":
            self._connect_peers()
        elif mode == "nccl":
            n = solver.halo_len
            self.to_left, self.to_right = make_buffer(n), make_buffer(n)
            self.from_left, self.from_right = make_buffer(n), make_buffer(n)
        else:
            raise ValueError("mode must be 'peer' or 'nccl'")

    def _connect_peers(self):
        """Exchange IPC handles of the three state buffers + flag words and map both neighbours."""
        s, dist = self.solver, self.dist
        if self.world == 1:
            mine = s.local_ptrs()
            s.connect_peers(mine, mine)
            return
        blobs = [None] * self.world
        dist.all_gather_object(blobs, s.ipc_export())
        opened = {}
        for r in {self.left, self.right}:
            opened[r] = s.ipc_open(blobs[r])
        s.connect_peers(opened[self.left], opened[self.right])
        self.barrier()

    def barrier(self):
        """Device + process barrier: required after uploads in peer mode (a neighbour may otherwise
        still be reading the halo columns the next step's first push overwrites)."""
        self.solver.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def check(self):
        """Peer mode: raise if an x sweep ever gave up waiting for a neighbour's halo columns (the kernels
        bound that wait so that a lost peer cannot hang the GPU; the state is invalid afterwards)."""
        if self.mode == "peer":
            self.solver.synchronize()  # pmw_synchronize reports the watchdog as an error

    # -- halo exchange ---------------------------------------------------------------------
    def exchange_halo_x(self, buf: int):
        s, dist = self.solver, self.dist
        s.pack_halo_x(buf, self.to_left.data_ptr(), self.to_right.data_ptr())
        if self.world == 1:
            # ring of one: my own edge columns come back as my halos (the reference's set_bc_x)
            self.from_left.copy_(self.to_right)
            self.from_right.copy_(self.to_left)
        else:
            ops = [dist.P2POp(dist.isend, self.to_left, self.left),
                   dist.P2POp(dist.isend, self.to_right, self.right),
                   dist.P2POp(dist.irecv, self.from_left, self.left),
                   dist.P2POp(dist.irecv, self.from_right, self.right)]
            if self.world == 2:
                # both neighbours are the same peer: order the pairs by tag-free FIFO semantics --
                # rank 0 sends (to_left, to_right), peer receives (from_right, from_left)
                ops = [dist.P2POp(dist.isend, self.to_left, self.left),
                       dist.P2POp(dist.irecv, self.from_right, self.right),
                       dist.P2POp(dist.isend, self.to_right, self.right),
                       dist.P2POp(dist.irecv, self.from_left, self.left)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        s.unpack_halo_x(buf, self.from_left.data_ptr(), self.from_right.data_ptr())
        self.exchanges += 1

    # -- time stepping ------------------------------------------------------------------------
    def evolve(self, nsteps: int = 1, dt: float | None = None):
        """step.py:85-143 on a slab: same stage sequence as pmw_evolve, plus the exchanges."""
        s = self.solver
        if self.mode == "peer":
            s.evolve(nsteps, dt)  # the exchange lives inside the stage kernels
            return
        for _ in range(nsteps):
            rev = s.reverse_direction
            for d in ((PMW_DIR_X, PMW_DIR_Z) if rev else (PMW_DIR_Z, PMW_DIR_X)):
                for rk in (1, 2, 3):
                    if d == PMW_DIR_X:
                        self.exchange_halo_x(PMW_BUF_STATE if rk == 1 else PMW_BUF_TMP)
                    s.evolve_stage(d, rk, dt)
            s.reverse_direction = not rev

    def stats(self):
        """Global (mass, energy): local reduction kernels + all-reduce of 2 doubles."""
        self.solver.stats_device(PMW_BUF_STATE, self.stats_buf.data_ptr())
        self.check()
        if self.world > 1:
            self.dist.all_reduce(self.stats_buf)
        out = self.stats_buf.cpu().numpy()
        return float(out[0]), float(out[1])
