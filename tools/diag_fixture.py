"""Diagnostic: error of the fused MLL kernel against a reference fixture (which assertion is tight and by how much)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L
dev = "cuda:0"
for case in sys.argv[1:] or ["lattice_d2_n1024_a2"]:
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", case + ".npz")))
    fam = 0 if str(g["family"]) == "lattice" else 1
    d, n, alpha = int(g["d"]), int(g["n"]), int(g["alpha"])
    t = int(g["t"]) if fam else 0
    xpts = torch.from_numpy(g["x"] if fam == 0 else g["xb"]).to(dev)
    yt = torch.from_numpy(g["ytilde"]).to(dev)
    ysq = (yt.abs() ** 2).reshape(1, n).contiguous()
    scale = torch.from_numpy(g["scale0"]).to(dev); ls = torch.from_numpy(g["lengthscales0"]).to(dev).reshape(1, d).contiguous(); noise = torch.from_numpy(g["noise0"]).to(dev)
    for z in ([None] + ([[int(v) for v in g["z"]]] if fam == 0 else [])):
        out, lam = L.mll_grad(fam, xpts, [alpha] * d, t, ysq, scale, ls, noise, want_grad=True, want_lam=True, z=z)
        out = out.cpu().numpy()[0]; lam = lam.cpu().numpy()[0]
        lam_ref = np.sqrt(n) * g["lam0"] + g["noise0"]
        print(case, "gen" if z else "x", "lam relmax(norm-wise) %.2e" % (np.abs(lam - lam_ref).max() / np.abs(lam_ref).max()),
              "lam rel elementwise max %.2e" % (np.abs(lam - lam_ref) / np.abs(lam_ref)).max(), "min|lam| %.2e" % np.abs(lam_ref).min(),
              "norm rel %.2e" % (abs(out[0] - g["norm_term0"].item()) / abs(g["norm_term0"].item())),
              "logdet rel %.2e" % (abs(out[1] - g["logdet0"].item()) / abs(g["logdet0"].item())),
              "gscale rel %.2e" % (abs(out[3] * g["scale0"][0] - g["grad_raw_scale0"][0]) / abs(g["grad_raw_scale0"][0])),
              "gls rel %.2e" % (np.abs(out[4:4 + d] * g["lengthscales0"] - g["grad_raw_lengthscales0"]).max() / np.abs(g["grad_raw_lengthscales0"]).max()))
