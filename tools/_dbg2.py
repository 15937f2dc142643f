import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L
dev = "cuda:0"
d, m = int(sys.argv[1]), int(sys.argv[2])
nz = float(sys.argv[3])
n = 1 << m
z = ([1, 182667, 469891, 498753, 110745, 446247, 250185, 118627] * 2)[:d]
xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
y = torch.cos(2 * np.pi * xp).sum(1) + 0.3 * torch.sin(2 * np.pi * xp[:, 0] * 3)
ysq = (L.fftbr(y).abs() ** 2).reshape(1, n)
scale = torch.full((1,), 2.5, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), nz, device=dev)
o_full, lam = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise, want_lam=True)
o_hs, _ = L.mll_grad(0, xp, [2] * d, 0, ysq, scale, ls, noise)
# independent: torch.fft on the natural-order lattice sequence
rs = scale.clone().requires_grad_(True); rl = ls[0].clone().requires_grad_(True)
j = torch.arange(n, device=dev)
zz = torch.tensor(z, device=dev)
delta = ((j[:, None] * zz[None, :]) % n).double() / n
parts = -(2 * np.pi) ** 4 / 24 * (delta ** 4 - 2 * delta ** 3 + delta ** 2 - 1 / 30)
k1 = rs * (1 + rl * parts).prod(-1)
lamt = torch.fft.fft(k1).real + nz
# natural frequency order: ysq is in the same (natural) order as the output of fftbr
loss = 0.5 * ((ysq[0] / lamt).sum() + torch.log(lamt).sum())
loss.backward()
print("full ", o_full[0].cpu().numpy())
print("hs   ", o_hs[0].cpu().numpy())
print("torch", np.array([float((ysq[0] / lamt).sum()), float(torch.log(lamt).sum()), 0.0, float(rs.grad)] + rl.grad.cpu().tolist()))
print("lam min", float(lamt.min()), "max", float(lamt.max()), "lam vs torch rel", float((lam[0].real - lamt).abs().max() / lamt.abs().max()), "min-bin rel", float(((lam[0].real - lamt).abs() / lamt.abs()).max()))
