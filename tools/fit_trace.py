"""Where the wall time of one fit(iterations=K) goes once the pooled context is warm: host-side phase timings (with and without
synchronisation points) by wrapping the loop's methods, and the kernel list of the same call from torch.profiler.
    python tools/fit_trace.py [K]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp  # noqa: E402
from fastgaussianprocesses_b200 import fast_gp as F  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
d, n, K = 8, 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 20


def mk():
    gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=bench.gen_vec(d)), device=dev)
    gp.get_x_next(n)
    return gp


gp0 = mk()
y = bench.f_synth(gp0.get_x(0, n)).contiguous()
for _ in range(3):
    g = mk()
    g.add_y_next(y)
    g.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1)

marks = []
SYNC = [False]


def wrap(obj, name):
    fn = getattr(obj, name)

    def w(*a, **k):
        if SYNC[0]:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        if SYNC[0]:
            torch.cuda.synchronize()
        marks.append((name, (time.perf_counter() - t0) * 1e6, t0))
        return r
    setattr(obj, name, w)


for nm in ("begin", "replay", "snapshot", "wait_snapshot", "finish", "close", "__init__"):
    wrap(F._FusedFitLoop, nm)
for nm in ("get_ytilde", "_get_ysq", "_fit_fused", "_new_fused_loop"):
    wrap(F.AbstractFastGP, nm)
wrap(F._FusedFitLoop, "eligible")
wrap(F._FitContext, "acquire")

for sync in (False, True):
    SYNC[0] = sync
    for rep in range(2):
        g = mk()
        g.add_y_next(y)
        torch.cuda.synchronize()
        marks.clear()
        t0 = time.perf_counter()
        g.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1)
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) * 1e6
    if not sync:
        print("timeline (us after fit() was entered: start, duration) of the last call:", [(nm, round((ts - t0) * 1e6), round(us)) for nm, us, ts in sorted(marks, key=lambda m: m[2])])
    agg = {}
    for nm, us, _ in marks:
        agg.setdefault(nm, [0, 0.0])
        agg[nm][0] += 1
        agg[nm][1] += us
    print("sync points" if sync else "no sync points", "total %.0f us:" % tot, {k: "%dx %.0f us" % (v[0], v[1]) for k, v in agg.items()})

SYNC[0] = False
g = mk()
g.add_y_next(y)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    g.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
