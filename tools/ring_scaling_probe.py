"""Weak-scaling probe of the slab ring under torchrun: for a list of per-GPU slab shapes and tuning switches, ms/step of
the ring against the same slab on rank 0 alone.   torchrun --nproc-per-node N tools/ring_scaling_probe.py nx,nz[,key=val...] ..."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
from pyminiweather_b200.slab import SlabRing
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.current_stream()


def timed(s, ring, steps, collective):
    sync = (lambda: (dist.barrier(), torch.cuda.synchronize())) if collective else torch.cuda.synchronize
    step = ring.evolve if ring is not None else s.evolve
    step(10); sync()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync(); e0.record(stream); step(steps); e1.record(stream); sync()
        ms = e0.elapsed_time(e1)
        if collective:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        best = min(best, ms)
    return best / steps


for spec in sys.argv[1:]:
    parts = spec.split(",")
    nx, nz = int(parts[0]), int(parts[1])
    tune = {k: int(v) for k, v in (p.split("=") for p in parts[2:])}
    steps = max(20, int(4e9 / (nx * nz * 70)))
    _, prof = new_case(8, nz, "thermal")  # 1-D profiles only
    rng = np.random.default_rng(20260101 + rank)
    st = np.zeros((4, nz + 4, nx + 4))
    for v, amp in enumerate((1e-3, 1e-1, 1e-1, 1e-1)):
        st[v, 2:-2, 2:-2] = amp * rng.uniform(-1.0, 1.0, size=(nz, nx))
    dx = 1e4 / nz

    def make(periodic):
        s = DeviceSolver(nx, nz, dx, dx, dx / 500.0, device=lr, periodic_x=periodic)
        s.set_stream(stream.cuda_stream)
        s.set_hydrostatic(*[getattr(prof, k) for k in HYDRO]); s.set_tuning(**tune)
        s.upload(0, st); s.upload(1, st)
        return s
    s = make(False)
    ring = SlabRing(s, rank, world, lambda n: torch.zeros(n, dtype=torch.float64, device="cuda"), dist, "peer")
    tn = timed(s, ring, steps, True)
    ring.check(); s.close()
    t1 = None
    if rank == 0:
        s1 = make(True); t1 = timed(s1, None, steps, False); s1.close()
    dist.barrier()
    # a second single-GPU figure: the same slab as a ring of ONE mapped onto itself (all of the protocol, no NVLink),
    # measured on every rank at the same time (do the GPUs of the box disturb each other?)
    s2 = make(False)
    mine = s2.local_ptrs(); s2.connect_peers(mine, mine)
    t_self = timed(s2, None, steps, False); s2.close()
    ts = torch.tensor([t_self], dtype=torch.float64, device="cuda")
    allts = [torch.empty_like(ts) for _ in range(world)]
    dist.all_gather(allts, ts)
    if rank == 0:
        print(f"{nx}x{nz} per GPU {tune}: ring of {world} {tn*1e3:8.1f} us/step   alone {t1*1e3:8.1f} us/step   efficiency {t1/tn:.3f}   "
              f"self-ring on every rank simultaneously: {[round(float(x.item())*1e3, 1) for x in allts]}", flush=True)
dist.destroy_process_group()
