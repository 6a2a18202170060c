"""A few fused steps of a 2048x1024-per-GPU slab ring under torchrun (target of tools/ring_ncu.sh)."""
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
nx, nz, steps = 2048, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 6
p, case = new_case(nx, nz, "thermal")
s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, device=lr, periodic_x=False)
s.set_stream(torch.cuda.current_stream().cuda_stream)
s.set_hydrostatic(*[getattr(case, k) for k in HYDRO])
s.upload(0, case.state); s.upload(1, case.state)
ring = SlabRing(s, rank, world, lambda n: torch.zeros(n, dtype=torch.float64, device="cuda"), dist, "peer")
ring.evolve(steps)
torch.cuda.synchronize(); dist.barrier()
print("rank", rank, "done; watchdog:", s.peer_timed_out(), flush=True)
dist.destroy_process_group()
