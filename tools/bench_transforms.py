"""Stand-alone transform timings (CUDA events, CUDA-graph replays for the warm number, L2 flushed for the cold one)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgaussianprocesses_b200 import _lib as L
from microbench import timeit
dev = "cuda:0"
torch.set_default_dtype(torch.float64)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)
res = {"cfg": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}}
for m in (16, 20, 22, 24):
    n = 1 << m
    x = torch.randn(n, device=dev)
    tc, _ = timeit(lambda: L.fwht(x), flush=flush)
    tw, _ = timeit(lambda: L.fwht(x))
    res["fwht_2^%d" % m] = {"us_cold": round(tc * 1e6, 1), "us_warm": round(tw * 1e6, 1), "alg16n_GBs_cold": round(16 * n / tc / 1e9), "alg16n_GBs_warm": round(16 * n / tw / 1e9)}
    tc, _ = timeit(lambda: L.fftbr(x), flush=flush)
    tw, _ = timeit(lambda: L.fftbr(x))
    res["fft_r2c_2^%d" % m] = {"us_cold": round(tc * 1e6, 1), "us_warm": round(tw * 1e6, 1), "alg24n_GBs_cold": round(24 * n / tc / 1e9), "alg24n_GBs_warm": round(24 * n / tw / 1e9)}
xb = torch.randn(64, 1 << 20, device=dev)
t, _ = timeit(lambda: L.fwht(xb), reps=5, warm=2)
res["fwht_batch64_2^20"] = {"ms": round(t * 1e3, 3), "alg16n_GBs": round(16 * xb.numel() / t / 1e9)}
x26 = torch.randn(1 << 26, device=dev)
t, _ = timeit(lambda: L.fwht(x26), reps=5, warm=2)
res["fwht_2^26"] = {"us": round(t * 1e6, 1), "alg16n_GBs": round(16 * x26.numel() / t / 1e9)}
xf = torch.randn(64, 1 << 20, device=dev)
t, _ = timeit(lambda: L.fftbr(xf), reps=5, warm=2)
res["fft_r2c_batch64_2^20"] = {"ms": round(t * 1e3, 3), "alg24n_GBs": round(24 * xf.numel() / t / 1e9)}
print(json.dumps(res))
