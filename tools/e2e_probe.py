"""Where the end-to-end time of one fit job goes (host wall clock with synchronisation points, then a cProfile of the same job)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
import bench
dev = torch.device("cuda:0")
d, n, K = 8, 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 50
def job(y_host, sync_points=False):
    gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=bench.gen_vec(d)), device=dev)
    gp.get_x_next(n)
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    gp.add_y_next(y_host)
    if sync_points:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    data = gp.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1, store_loss_hist=True)
    if sync_points:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    hyp = [gp.scale.detach().cpu(), gp.lengthscales.detach().cpu(), data["loss_hist"].cpu()]
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [round((b - a) * 1e3, 3) for a, b in zip(t[:-1], t[1:])], round((t[-1] - t[0]) * 1e3, 3)
gp0 = fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=bench.gen_vec(d)), device=dev)
x = gp0.get_x_next(n)
y_host = bench.f_synth(x).cpu().pin_memory()
for _ in range(3):
    job(y_host)
print("phases ms [add_y_next, fit, results] total:", job(y_host, True), job(y_host, True))
print("no sync points:", job(y_host), job(y_host))
pr = cProfile.Profile()
pr.enable()
job(y_host)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
