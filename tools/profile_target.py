"""Small driver for ncu: runs each hot kernel a few times at the C3/C4 sizes."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgaussianprocesses_b200 import _lib as L

dev = "cuda:0"
torch.set_default_dtype(torch.float64)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d, n = 8, 1 << 20
z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]
xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
if which in ("all", "fft"):
    x = torch.randn(n, device=dev)
    for _ in range(reps):
        y = L.fftbr(x)
        x2 = L.ifftbr(y)
if which in ("all", "fwht"):
    x = torch.randn(1 << 24, device=dev)
    for _ in range(reps):
        y = L.fwht(x)
    x = torch.randn(1 << 20, device=dev)
    for _ in range(reps):
        y = L.fwht(x)
if which in ("all", "mll"):
    ysq = torch.rand(1, n, device=dev)
    scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
    for _ in range(reps):
        L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise)
if which in ("all", "mllz"):  # generator mode (what fit() runs): the points are regenerated from the index, half-spectrum mode
    ysq = torch.rand(1, n, device=dev)
    ysq = (ysq + ysq.flip(-1).roll(1, -1)) / 2
    scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
    for _ in range(reps):
        L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, z=z)
if which in ("all", "pmean"):
    xs = torch.rand(1 << 13, d, device=dev)
    co = torch.randn(1, n, device=dev)
    for _ in range(reps):
        L.post_mean(0, xs, xp, [2] * d, 0, 1.0, [0.5] * d, co)
if which in ("all", "pvar"):  # lattice post_var: cross-pair kernel -> in-place c2c transform -> pair reduction
    xs = torch.rand(256, d, device=dev)
    ysq = torch.rand(1, n, device=dev)
    scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
    _, lam = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, want_lam=True)
    for _ in range(reps):
        L.post_var(0, xs, xp, [2] * d, 0, 1.0, [0.5] * d, lam[0])
if which in ("all", "pvarz"):  # fused generator-form lattice post_var: pv_passA (kernel evaluation + block transform) -> pv_passB (+ pair reduction)
    xs = torch.rand(256, d, device=dev)
    ysq = torch.rand(1, n, device=dev)
    scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
    _, lam = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, want_lam=True)
    for _ in range(reps):
        L.post_var_z(xs, z, np.linspace(0.1, 0.9, d), n, [2] * d, 1.0, [0.5] * d, lam[0])
if which in ("netpm",):  # net post_mean d = 16 (configs[3] family), alpha = 2, t = 52: which pipe binds the inner loop?
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(16, seed=7, t=52), device=dev, noise=1e-8)
    xn = gp.get_x_next(n)
    gp.add_y_next(torch.cos(2 * np.pi * xn).sum(1))
    gp.coeffs
    xs = torch.rand(1 << 12, 16, device=dev)
    for _ in range(reps):
        gp.post_mean(xs)
if which in ("netpv",):  # fused net post_var d = 8
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(8, seed=7, t=52), device=dev, noise=1e-8)
    xn = gp.get_x_next(n)
    gp.add_y_next(torch.cos(2 * np.pi * xn).sum(1))
    xs = torch.rand(128, 8, device=dev)
    for _ in range(reps):
        gp.post_var(xs)
torch.cuda.synchronize()
print("ok")
