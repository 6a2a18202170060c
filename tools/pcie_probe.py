"""What the host link gives: flat pinned copies H2D / D2H alone and both at once, against the pitched (2-D) copies the
state upload / download use.   usage: python tools/pcie_probe.py [nx nz]"""
import sys, time
import torch
nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
n = 4 * (nz + 4) * (nx + 4)
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
pitch = (14 + nx + 8 + 15) // 16 * 16
dp = torch.empty(4 * (nz + 4) * pitch, dtype=torch.float64, device="cuda").view(4 * (nz + 4), pitch)
hv = h1.view(4 * (nz + 4), nx + 4)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nb = n * 8


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


a = t(lambda: d1.copy_(h1, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True)); c = t(both)
p = t(lambda: dp[:, 14:14 + nx + 4].copy_(hv, non_blocking=True)); q = t(lambda: hv.copy_(dp[:, 14:14 + nx + 4], non_blocking=True))
print(f"{nb/1e6:.1f} MB   flat H2D {nb/a/1e9:.1f} GB/s   flat D2H {nb/b/1e9:.1f} GB/s   both at once {2*nb/c/1e9:.1f} GB/s summed "
      f"({c*1e3:.3f} ms)   pitched H2D {nb/p/1e9:.1f}   pitched D2H {nb/q/1e9:.1f}")
