"""post_var points/s through the public API (lattice d=8 n=2^20 / 2^16, net d=4 n=2^16 / d=8 n=2^20)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
from microbench import timeit
dev = "cuda:0"
res = {}
for name, mk, n, M in (("lattice_d8_n2^20", lambda: fgp.FastGPLattice(fgp.Lattice(8, seed=7), device=dev), 1 << 20, 512),
                       ("lattice_d8_n2^16", lambda: fgp.FastGPLattice(fgp.Lattice(8, seed=7), device=dev), 1 << 16, 4097),
                       ("dnb2_d4_n2^16", lambda: fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(4, seed=7), device=dev), 1 << 16, 4096),
                       ("dnb2_d8_n2^20", lambda: fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(8, seed=7), device=dev, noise=1e-8), 1 << 20, 512)):
    gp = mk()
    x = gp.get_x_next(n)
    gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
    xs = torch.rand(M, x.shape[1], device=dev)
    gp.post_var(xs[:8])
    t, _ = timeit(lambda: gp.post_var(xs), reps=3, warm=1, graph=False)
    res["post_var_%s_m%d" % (name, M)] = {"ms": round(t * 1e3, 3), "pts_per_s": round(M / t)}
    del gp
print(json.dumps(res))
