"""x-slab ring over NCCL on real GPUs (needs >= 2 devices; skipped otherwise): the sharded run
must reproduce the single-GPU run of the same domain bit for bit, and the single-GPU run is tied
to the oracle elsewhere (test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import HYDRO, synthetic_case  # noqa: E402

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, nx, nz, nsteps, out_dir, mode):
    import torch
    import torch.distributed as dist
    from pyminiweather_b200.engine import DeviceSolver
    from pyminiweather_b200.slab import SlabRing
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _, whole = synthetic_case(nx, nz, seed=9)
        nxl = nx // world
        st = np.ascontiguousarray(whole.state[:, :, rank * nxl: (rank + 1) * nxl + 4])
        s = DeviceSolver(nxl, nz, whole.dx, whole.dz, whole.dt, device=rank, periodic_x=False)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        s.set_hydrostatic(*[getattr(whole, n) for n in HYDRO])
        s.upload(0, st); s.upload(1, st)
        ring = SlabRing(s, rank, world, lambda n: torch.zeros(n, dtype=torch.float64, device="cuda"), dist, mode)
        ring.evolve(nsteps)
        m, e = ring.stats()
        timed_out = s.peer_timed_out() if mode == "peer" else False
        dist.barrier()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=s.download(0), stats=np.array([m, e]),
                 timed_out=timed_out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("mode", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_ring_matches_single_gpu(world, mode, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from pyminiweather_b200.engine import DeviceSolver
    nx, nz, nsteps = 256 * world, 96, 5  # 5 steps: both sweep orders, incl. X-S3 -> X-S1
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, nx, nz, nsteps, str(tmp_path), mode), nprocs=world, join=True)
    _, whole = synthetic_case(nx, nz, seed=9)
    one = DeviceSolver(nx, nz, whole.dx, whole.dz, whole.dt)
    one.set_hydrostatic(*[getattr(whole, n) for n in HYDRO])
    one.upload(0, whole.state); one.upload(1, whole.state)
    one.evolve(nsteps)
    want = one.download(0)
    ws = one.stats(0)
    nxl = nx // world
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert not bool(got["timed_out"])
        assert np.array_equal(got["state"][:, 2:-2, 2:-2], want[:, 2:-2, 2 + r * nxl: 2 + (r + 1) * nxl])
        assert abs(got["stats"][0] - ws[0]) / ws[0] < 1e-13 and abs(got["stats"][1] - ws[1]) / ws[1] < 1e-13
    one.close()


def test_peer_ring_of_one_equals_periodic_single_gpu():
    """A ring of one slab mapped onto itself through the peer path == the periodic context."""
    from pyminiweather_b200.engine import DeviceSolver
    _, whole = synthetic_case(380, 50, seed=2)
    a = DeviceSolver(380, 50, whole.dx, whole.dz, whole.dt)
    b = DeviceSolver(380, 50, whole.dx, whole.dz, whole.dt, periodic_x=False)
    for s in (a, b):
        s.set_hydrostatic(*[getattr(whole, n) for n in HYDRO])
        s.upload(0, whole.state); s.upload(1, whole.state)
    mine = b.local_ptrs()
    b.connect_peers(mine, mine)
    a.evolve(5); b.evolve(5)
    assert not b.peer_timed_out()
    assert np.array_equal(a.download(0)[:, 2:-2, 2:-2], b.download(0)[:, 2:-2, 2:-2])
    a.close(); b.close()
