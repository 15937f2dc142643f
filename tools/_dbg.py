import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L
dev = "cuda:0"
for d, m in [(4, 16), (5, 16), (6, 16), (7, 16), (8, 16), (3, 16), (2, 16), (16, 16)]:
    n = 1 << m
    z = ([1, 182667, 469891, 498753, 110745, 446247, 250185, 118627] * 2)[:d]
    xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    ysq = torch.rand(1, n, device=dev, generator=g)
    if os.environ.get("REALY") == "1":
        y = torch.cos(2 * np.pi * xp).sum(1) + 0.3 * torch.sin(2 * np.pi * xp[:, 0] * 3)
        ysq = (L.fftbr(y).abs() ** 2).reshape(1, n)
    else:
        ysq = (ysq + ysq.flip(-1).roll(1, -1)) / 2  # even in the natural frequency order: ysq[k] = ysq[n-k]
    scale = torch.full((1,), 2.5, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), float(os.environ.get("NOISE", "1e-6")), device=dev)
    o_full, lam = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, want_lam=True)
    o_hs, _ = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise)
    o_z, _ = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, z=z)
    print(d, m, "x-mode rel", ((o_hs - o_full).abs() / o_full.abs()).cpu().numpy().round(12), "\n   z-mode rel", ((o_z - o_full).abs() / o_full.abs()).cpu().numpy().round(12))
