"""Timing helper shared by the sweep tools (development tool)."""
def bench(s, d, reps=40):
    for _ in range(5):
        for r in (1, 2, 3): s.evolve_stage(d, r)
    s.synchronize()
    res = []
    for rk in (1, 2, 3):
        tot = 0.0; n = 0
        for _ in range(reps):
            for r in (1, 2, 3):
                if r == rk: s.stage_timing(True)
                s.evolve_stage(d, r)
                if r == rk:
                    ms, k = s.stage_timing_read(); s.stage_timing(False); tot += ms * k; n += k
        res.append(tot / n * 1e3)
    return res
