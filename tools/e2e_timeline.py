"""GPU timeline of one end-to-end fit job (host y -> add_y_next -> fit(K) -> results on the host) from torch.profiler: when the H2D copy,
the first / last fit kernel and the D2H copies run relative to the host's start, and how long the GPU idles in between.
    python tools/e2e_timeline.py [K]"""
import os, sys, time
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
import bench
dev = torch.device("cuda:0")
d, n, K = 8, 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 20
mk = lambda: fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=bench.gen_vec(d)), device=dev)


def job(gp, y_host):
    t0 = time.perf_counter()
    gp.add_y_next(y_host)
    t1 = time.perf_counter()
    data = gp.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1, store_loss_hist=True)
    t2 = time.perf_counter()
    hyp = [gp.scale.detach().cpu(), gp.lengthscales.detach().cpu(), data["loss_hist"].cpu()]
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return [round((b - a) * 1e6) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t0, t3))]


gp0 = mk()
x = gp0.get_x_next(n)
y_host = bench.f_synth(x).cpu().pin_memory()
for _ in range(3):
    g = mk(); g.get_x_next(n); job(g, y_host)
outs = []
for _ in range(5):
    g = mk(); g.get_x_next(n); torch.cuda.synchronize(); outs.append(job(g, y_host))
print("host us [add_y_next, fit, results, total] x5:", outs)
g = mk(); g.get_x_next(n); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    job(g, y_host)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
busy, last_end, rows = 0.0, None, []
for e in evs:
    s, en = e.time_range.start - t0, e.time_range.end - t0
    gap = 0 if last_end is None else s - last_end
    rows.append((round(s, 1), round(en - s, 1), round(gap, 1), e.name[:60]))
    busy += en - s
    last_end = max(en, last_end or 0)
print("GPU span %.0f us, busy %.0f us, events %d" % (last_end, busy, len(evs)))
big = [r for r in rows if r[2] > 8 or r[1] > 60]
print("events with a gap > 8 us before them or longer than 60 us (start, dur, gap_before, name):")
for r in big[:40]:
    print("  ", r)
print("first 14 events:")
for r in rows[:14]:
    print("  ", r)
print("last 8 events:")
for r in rows[-8:]:
    print("  ", r)
