"""Kernel-level timings on one B200 (CUDA events, L2 flushed between timed launches where noted)."""
import json
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgaussianprocesses_b200 import _lib as L

dev = "cuda:0"
torch.set_default_dtype(torch.float64)


def timeit(fn, reps=20, warm=3, flush=None, graph=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if graph and flush is None:
        # capture `inner` back-to-back calls in a CUDA graph so that Python/ctypes launch overhead is not timed
        inner = 10
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=st):
                for _ in range(inner):
                    fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, reps // 4)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / inner)
        return float(np.median(ts)), float(np.min(ts))
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    res = {}
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)
    # fp64 peak
    iters = 4000
    fl = L.fp64_peak_probe(iters, dev)
    t, tmin = timeit(lambda: L.fp64_peak_probe(iters, dev), reps=5)
    res["fp64_tflops"] = fl / tmin / 1e12
    # copy bandwidth
    a = torch.empty(2 ** 28, device=dev); b = torch.empty_like(a)
    t, tmin = timeit(lambda: b.copy_(a), reps=5)
    res["copy_gbs"] = 2 * a.numel() * 8 / tmin / 1e9
    del a, b
    for m in (16, 20, 22, 24):
        n = 1 << m
        x = torch.randn(n, device=dev)
        t, tmin = timeit(lambda: L.fftbr(x), flush=flush)
        t2, tmin2 = timeit(lambda: L.fftbr(x))
        res["fft_r2c_2^%d" % m] = {"us_cold": t * 1e6, "us_warm": t2 * 1e6, "GBs_alg24n_cold": 24 * n / t / 1e9, "GBs_alg24n_warm": 24 * n / t2 / 1e9}
        if m <= 26:
            t, tmin = timeit(lambda: L.fwht(x), flush=flush)
            t2, tmin2 = timeit(lambda: L.fwht(x))
            res["fwht_2^%d" % m] = {"us_cold": t * 1e6, "us_warm": t2 * 1e6, "GBs_alg16n_cold": 16 * n / t / 1e9, "GBs_alg16n_warm": 16 * n / t2 / 1e9}
    x = torch.randn(1 << 26, device=dev)
    t, _ = timeit(lambda: L.fwht(x), reps=5)
    res["fwht_2^26"] = {"us": t * 1e6, "GBs_alg16n": 16 * x.numel() / t / 1e9}
    del x
    # MLL iteration
    for fam, d, m in ((0, 8, 20), (0, 8, 18), (1, 4, 16), (1, 16, 24), (0, 2, 10)):
        n = 1 << m
        if fam == 0:
            z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627][:d]
            xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
            t_ = 0
        else:
            from scipy.stats import qmc
            t_ = 52
            C = torch.from_numpy((qmc.Sobol(d, scramble=False, bits=32)._sv.astype(np.uint64) << np.uint64(20)).astype(np.int64)).to(dev)
            xp, _ = L.dnb2_points(C, list(range(1, d + 1)), t_, 0, n)
        ysq = torch.rand(1, n, device=dev)
        scale = torch.ones(1, device=dev); ls = torch.full((1, d), 0.5, device=dev); noise = torch.full((1,), 1e-6, device=dev)
        f = lambda: L.mll_grad(fam, xp, [2] * d, t_, ysq, scale, ls, noise)
        t, tmin = timeit(f, reps=20)
        tc, _ = timeit(f, reps=10, flush=flush)
        res["mll_grad_fam%d_d%d_2^%d" % (fam, d, m)] = {"us_warm": t * 1e6, "us_cold": tc * 1e6, "iters_per_s_warm": 1 / t}
        del xp, ysq
    # post_mean
    for fam, d, m, mt in ((0, 8, 20, 14), (0, 8, 20, 17), (1, 16, 20, 14), (0, 2, 10, 12), (1, 4, 16, 14)):
        n = 1 << m
        M = 1 << mt
        if fam == 0:
            z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627][:d]
            xp = L.lattice_points(z, np.linspace(0.1, 0.9, d), 0, n, dev)
            t_ = 0
        else:
            from scipy.stats import qmc
            t_ = 52
            C = torch.from_numpy((qmc.Sobol(d, scramble=False, bits=32)._sv.astype(np.uint64) << np.uint64(20)).astype(np.int64)).to(dev)
            xp, _ = L.dnb2_points(C, list(range(1, d + 1)), t_, 0, n)
        xs = torch.rand(M, d, device=dev)
        co = torch.randn(1, n, device=dev)
        f = lambda: L.post_mean(fam, xs, xp, [2] * d, t_, 1.0, [0.5] * d, co)
        t, tmin = timeit(f, reps=5, warm=2)
        slots = (5 * d + 1) * float(M) * n
        res["post_mean_fam%d_d%d_n2^%d_m2^%d" % (fam, d, m, mt)] = {"ms": t * 1e3, "pts_per_s": M / t, "fp64_slot_TFLOPs_equiv": 2 * slots / t / 1e12}
    # batched transforms (the shape post_var and batched fits use): 64 rows of 2^20 through the two-pass pipeline
    xb_ = torch.randn(64, 1 << 20, device=dev)
    t, _ = timeit(lambda: L.fwht(xb_), reps=5, warm=2)
    res["fwht_batch64_2^20"] = {"ms": t * 1e3, "GBs_alg16n": 16 * xb_.numel() / t / 1e9}
    t, _ = timeit(lambda: L.fftbr(xb_), reps=5, warm=2)
    res["fft_r2c_batch64_2^20"] = {"ms": t * 1e3, "GBs_alg24n": 24 * xb_.numel() / t / 1e9}
    del xb_
    # post_var and batched fits through the public API
    import fastgaussianprocesses_b200 as fgp
    for name, mk, n, M in (("lattice_d8_n2^20", lambda: fgp.FastGPLattice(fgp.Lattice(8, seed=7), device=dev), 1 << 20, 256),
                           ("dnb2_d4_n2^16", lambda: fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(4, seed=7), device=dev), 1 << 16, 4096)):
        gp = mk()
        x = gp.get_x_next(n)
        gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
        xs = torch.rand(M, x.shape[1], device=dev)
        gp.post_var(xs[:8])
        t, _ = timeit(lambda: gp.post_var(xs), reps=3, warm=1, graph=False)
        res["post_var_%s_m%d" % (name, M)] = {"ms": t * 1e3, "pts_per_s": M / t}
        del gp
    # C5: 64 independent lattice GPs, d=8, n=2^18, one object with shape_batch=[64] and per-GP hyperparameters
    Bt, d, n = 64, 8, 1 << 18
    gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7), device=dev, shape_batch=[Bt], scale=torch.ones(Bt, 1), lengthscales=torch.ones(Bt, d), noise=torch.full((Bt, 1), 1e-8))
    x = gp.get_x_next(n)
    fr = torch.arange(1, Bt + 1, device=dev, dtype=torch.float64)[:, None]
    gp.add_y_next(torch.cos(2 * np.pi * x.sum(1)[None, :] * fr / 8))
    st = gp.fit_stepper()
    for _ in range(3):
        st.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        st.step()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    res["batched_fit_64x_lattice_d8_2^18"] = {"ms_per_batched_iteration": t * 1e3, "gp_iterations_per_s": Bt / t}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
