"""Timing probe of the persistent fit kernel (mll_coop_kernel) against the three-launch route.

    python tools/coop_probe.py [--log2n 20] [--d 8] [--family lattice] [--B 1] [--iters 200]

Prints one JSON line: microseconds per fit iteration for (a) one iteration per launch, eager; (b) `chunk` iterations per
launch; (c) the three-launch route (FGP_COOP=0), eager and from a CUDA graph.  Geometry / grid overrides come from the
environment (FGP_CAP_C, FGP_COLS_LOG2, FGP_COOP_CTAS, ...; tools/tune_mll.py sweeps them in subprocesses).
With a -DFGP_TIMING build (FGP_B200_LIB=...) it also prints the phase stamps of the last launch."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp  # noqa: E402
from fastgaussianprocesses_b200 import _lib as L  # noqa: E402

GEN_VEC = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]


def make(family, d, n, B, dev):
    kw = {}
    if B > 1:
        kw = dict(shape_batch=torch.Size([B]), shape_lengthscales=torch.Size([B, d]), shape_scale=torch.Size([B, 1]))
    if family == "lattice":
        z = GEN_VEC[:d] if d <= 8 else None
        gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7, generating_vector=np.asarray(z, dtype=np.uint64) if z else None), device=dev, **kw)
    else:
        gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=7), device=dev, **kw)
    x = gp.get_x_next(n)
    j = torch.arange(1, d + 1, device=dev, dtype=x.dtype)
    y = torch.cos(2 * np.pi * x).mul(1.0 / j).sum(1) + torch.sin(2 * np.pi * x[:, 0]) * torch.cos(2 * np.pi * x[:, -1])
    if B > 1:
        y = torch.stack([y * (1 + 0.1 * k) for k in range(B)])
    gp.add_y_next(y)
    return gp


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--family", default="lattice")
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--skip-three", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = 1 << args.log2n
    out = {"log2n": args.log2n, "d": args.d, "family": args.family, "B": args.B,
           "env": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}}
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)
    os.environ["FGP_COOP"] = "1"
    gp = make(args.family, args.d, n, args.B, dev)
    st = gp.fit_stepper()
    out["multi"] = bool(st.multi)
    for _ in range(5):
        st.step()
    out["coop_1_per_launch_us"] = timed(st.step, args.iters)
    # L2 flushed before every step, each step timed alone
    ts = []
    for _ in range(20):
        flush.zero_()
        ts.append(timed(st.step, 1))
    out["coop_1_per_launch_cold_us"] = float(np.median(ts))
    if st.multi:
        st.replay(args.chunk)
        out["coop_chunk_us_per_iter"] = timed(lambda: st.replay(args.chunk), max(1, args.iters // args.chunk)) / args.chunk
        # the same from a CUDA graph of single-iteration launches (cooperative launches are capturable)
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g.capture_begin()
                for _ in range(16):
                    L.fit_iterations(st.problem, st.layout, 1)
                g.capture_end()
            torch.cuda.current_stream().wait_stream(s)
            g.replay()
            out["coop_graph16_us_per_iter"] = timed(g.replay, max(1, args.iters // 16)) / 16
        except Exception as e:  # noqa: BLE001
            out["coop_graph16_error"] = str(e)[:200]
    try:
        stamps_fn = ctypes.CDLL(L.LIB_PATH).fgp_debug_stamps
    except AttributeError:
        stamps_fn = None
    if stamps_fn is not None and st.multi:
        st.step()
        torch.cuda.synchronize()
        buf = np.zeros((1024, 48), dtype=np.int64)
        stamps_fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
        stamps_fn(buf.ctypes.data, 1024)
        live = buf[buf[:, 0] != 0]
        t0 = live[:, 0].min()
        rel = (live[:, :7] - t0) * 1e-3  # microseconds since the first CTA started
        names = ["start", "A done", "bar1 passed", "B done", "bar2 passed", "C done", "exit"]
        out["stamps_ctas"] = int(len(live))
        out["stamps_us"] = {nm: {"min": float(rel[:, k].min()), "mean": float(rel[:, k].mean()), "max": float(rel[:, k].max())} for k, nm in enumerate(names)}
        ck = (live[:, 8:15] - live[:, 8:9]) / 1.965e3  # per-CTA clock64 cycles -> microseconds at 1965 MHz
        rel = ck + rel[:, 0:1]
        out["stamps_clock_us"] = {nm: {"min": float(rel[:, k].min()), "mean": float(rel[:, k].mean()), "max": float(rel[:, k].max())} for k, nm in enumerate(names)}
        dur = {"A": rel[:, 1] - rel[:, 0], "wait1": rel[:, 2] - rel[:, 1], "B": rel[:, 3] - rel[:, 2], "wait2": rel[:, 4] - rel[:, 3], "C": rel[:, 5] - rel[:, 4]}
        out["phase_us"] = {k: {"min": float(v.min()), "mean": float(v.mean()), "max": float(v.max())} for k, v in dur.items()}
    st.close()
    if not args.skip_three:
        os.environ["FGP_COOP"] = "0"
        gp0 = make(args.family, args.d, n, args.B, dev)
        st0 = gp0.fit_stepper()
        for _ in range(5):
            st0.step()
        out["three_launch_graph_us"] = timed(st0.step, args.iters)
        ts = []
        for _ in range(20):
            flush.zero_()
            ts.append(timed(st0.step, 1))
        out["three_launch_cold_us"] = float(np.median(ts))
        st0.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
