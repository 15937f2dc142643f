"""Phase timing of mll_passA from clock64 stamps.  Needs a -DFGP_TIMING build:
    FGP_LIB_DIR=/tmp/lib_dbg FGP_BUILD_DEFS=-DFGP_TIMING python -m fastgaussianprocesses_b200.build
    FGP_B200_LIB=/tmp/lib_dbg/libfgp_b200.so python tools/timing_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L
dev = "cuda:0"
for m in (18, 20):
    n, d = 1 << m, 8
    z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]
    xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
    ysq = torch.rand(1, n, device=dev); one = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); nz = torch.full((1,), 1e-6, device=dev)
    for rep in range(3):
        # want_grad=0: pass C does not run, so partC keeps pass A's debug stamps
        out, _ = L.mll_grad(0, xp, [2] * d, 0, ysq, one, ls, nz, want_grad=False, z=z)
    torch.cuda.synchronize()
    ws = L._workspaces[("mll", 0)]
    nbytes = L.load().fgp_mll_workspace_bytes(0, n, d, 1)
    ctasA = n >> 12
    # partC offset: wbytes + pb
    al = lambda v: (v + 255) & ~255
    off = al(n * 16) + al((1 << 12) // 8 * 3 * 8)
    t = ws.view(torch.uint8)[off:off + ctasA * 32].view(torch.float64).reshape(ctasA, 4).cpu().numpy()
    fill, start, total, sm = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    print("n=2^%d ctas=%d: fill cycles mean %.0f (min %.0f max %.0f); whole CTA after hyp: mean %.0f (min %.0f max %.0f); start spread %.0f cycles; CTAs per SM max %d"
          % (m, ctasA, fill.mean(), fill.min(), fill.max(), total.mean(), total.min(), total.max(), start.max() - start.min(), np.bincount(sm.astype(int)).max()))
