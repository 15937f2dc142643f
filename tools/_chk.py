import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L
dev = "cuda:0"; d = 8; n = 1 << 20
z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]
xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
xs = torch.rand(256, d, device=dev)
ysq = torch.rand(1, n, device=dev)
scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
_, lam = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, want_lam=True)
c0 = L.launch_count()
pv = L.post_var(0, xs, xp, [2] * d, 0, 1.0, [0.5] * d, lam[0])
torch.cuda.synchronize()
print("launches", L.launch_count() - c0, pv[:4].tolist())
