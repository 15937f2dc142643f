"""Digital-net timings (BASELINE.json configs[1] / configs[3] shapes): post_mean points/s, fit iterations, batched FWHT.
    python tools/bench_net.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp  # noqa: E402

dev = "cuda:0"


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


res = {}
for d, m, M, t in ((16, 20, 1 << 13, 52), (4, 16, 1 << 16, 52), (16, 20, 1 << 12, 63)):
    gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=7, t=t), device=dev, noise=1e-8)
    x = gp.get_x_next(1 << m)
    gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
    gp.coeffs
    xs = torch.rand(M, d, device=dev)
    tm = timed(lambda: gp.post_mean(xs), 3)
    res["post_mean_dnb2_d%d_n2^%d_t%d" % (d, m, t)] = {"ms": round(tm * 1e3, 3), "pts_per_s": round(M / tm), "T_pair_dims_per_s": round(M * (1 << m) * d / tm / 1e12, 3)}
    K = 30
    tf = timed(lambda: gp.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1), 2)
    res["fit_dnb2_d%d_n2^%d_t%d" % (d, m, t)] = {"us_per_iteration": round(tf / (K + 1) * 1e6, 1)}
    del gp
print(json.dumps(res))
