"""Batched stand-alone FWHT: two launches (pass A of all items, pass B of all items) against the fused persistent kernel."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgaussianprocesses_b200 import _lib as L
from microbench import timeit
dev = "cuda:0"
torch.set_default_dtype(torch.float64)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)
res = {"cfg": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}}
for B, m in ((64, 20), (16, 22), (256, 18), (1024, 16), (4, 24), (1, 22), (1, 20), (1, 24)):
    n = 1 << m
    xb = torch.randn(B, n, device=dev)
    ref = L.fwht(xb, fused=False)
    assert torch.equal(L.fwht(xb, fused=True), ref)
    r = {}
    for name, fused in (("two_pass", False), ("fused", True)):
        t, _ = timeit(lambda: L.fwht(xb, fused=fused), reps=5, warm=2, flush=flush if B * n * 8 < (200 << 20) else None)
        r[name] = {"ms": round(t * 1e3, 4), "alg16n_GBs": round(16 * xb.numel() / t / 1e9)}
    res["fwht_%dx2^%d" % (B, m)] = r
print(json.dumps(res))
