"""Batched FWHT: two launches (wht_passA + wht_passB) against the fused persistent kernel (TMA-staged unless FGP_FUSED_NO_TMA=1).
    python tools/bench_fwht_fused.py        -> one JSON line; 16 n bytes per item are the algorithmic bytes"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from fastgaussianprocesses_b200 import _lib as L  # noqa: E402

dev = "cuda:0"
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}}
for B, m in ((64, 20), (256, 18), (16, 22), (1024, 16), (8, 20)):
    x = torch.randn(B, 1 << m, device=dev)
    row = {}
    ref = L.fwht(x, fused=False)
    got = L.fwht(x, fused=True)
    row["max_abs_diff"] = float((ref - got).abs().max())
    for name, fused in (("two_launch", False), ("fused", True)):
        for _ in range(3):
            L.fwht(x, fused=fused)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            L.fwht(x, fused=fused)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        row[name] = {"ms": round(t * 1e3, 4), "alg16n_GBs": round(16 * x.numel() / t / 1e9)}
    res["%dx2^%d" % (B, m)] = row
    del x, ref, got
print(json.dumps(res))
