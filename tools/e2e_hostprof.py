"""Host-side profile (cProfile, microseconds) of add_y_next(host y) and of fit(K) on a warm pooled context.  python tools/e2e_hostprof.py [K]"""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
import bench
dev = torch.device("cuda:0")
d, n, K = 8, 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 20
mk = lambda: fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=bench.gen_vec(d)), device=dev)
gp0 = mk()
x = gp0.get_x_next(n)
y_host = bench.f_synth(x).cpu().pin_memory()
for _ in range(3):
    g = mk(); g.get_x_next(n); g.add_y_next(y_host); g.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1, store_loss_hist=True)


def show(pr, title, top=22):
    st = pstats.Stats(pr).stats
    rows = sorted(((tt * 1e6, ct * 1e6, nc, "%s:%d(%s)" % (os.path.basename(f), l, fn)) for (f, l, fn), (cc, nc, tt, ct, callers) in st.items()), reverse=True)
    print(title, "-- tottime us, cumtime us, calls, function")
    for r in rows[:top]:
        print("   %7.1f %8.1f %4d  %s" % r)


for what in ("add_y_next", "fit"):
    g = mk(); g.get_x_next(n); torch.cuda.synchronize()
    pr = cProfile.Profile()
    if what == "add_y_next":
        pr.enable(); g.add_y_next(y_host); pr.disable()
    else:
        g.add_y_next(y_host); torch.cuda.synchronize()
        pr.enable(); g.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1, store_loss_hist=True); pr.disable()
    torch.cuda.synchronize()
    show(pr, what)
