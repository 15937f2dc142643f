#!/usr/bin/env python
"""bench.py -- fit() MLL+gradient iterations/s and post_mean points/s, FastGPLattice d=8 n=2^20 FP64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 20] [--d 8]

A "step" is one fit() iteration: the fused eigen-solve (first kernel column -> FFT -> log-det + quadratic form ->
inverse FFT -> hyperparameter gradients) followed by the Rprop update.  Three measurements per run:
  value     device-resident: K steps, each timed alone with CUDA events, L2 flushed between steps (untimed);
  e2e       the public API with HOST buffers: add_y_next(host y) -> fit(iterations=K) -> hyperparameters back on the
            host, host<->device copies inside the timed region;
  post_mean points/s of post_mean on m test points per GPU (device-resident, and e2e with host x* / host result).
N>1 (torchrun): independent GPs per rank (batched fits / restarts: no data-path collective) and test points sharded
across ranks with one NCCL all_gather of the results; max over ranks, weak scaling.
`--impl reference` times the reference algorithm's CPU port (oracle/, torch float64 on the host cores).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)

# dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, n=2^20 d=8 (profiles/README.md, snapshot g)
NCU_TRAFFIC = {("mll_passA", 20, 8): 159232, ("mll_passB", 20, 8): 8623616, ("mll_passC", 20, 8): 4380416}

GEN_VEC = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]


def f_synth(x):
    """Smooth periodic test function (multitask/fgp_lattice.ipynb cell-4 style)."""
    j = torch.arange(1, x.shape[1] + 1, device=x.device, dtype=x.dtype)
    return torch.cos(2 * math.pi * x).mul(1.0 / j).sum(1) + torch.sin(2 * math.pi * x[:, 0]) * torch.cos(2 * math.pi * x[:, -1])


def gen_vec(d):
    if d <= len(GEN_VEC):
        return np.asarray(GEN_VEC[:d], dtype=np.uint64)
    from fastgaussianprocesses_b200.sequences import default_generating_vector
    return default_generating_vector(d)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md clocks line)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            allsm = []
            for ts, line in self.rows:
                try:
                    allsm.append((float(line.split(",")[0]), float(line.split(",")[1])))
                except Exception:
                    pass
            sm = [a for a, _ in allsm]
            mx = [b for _, b in allsm]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(n, d, iters, pm_points, seed=7):
    """The reference algorithm's CPU port (oracle/) on this box's host cores: fit iterations/s and post_mean points/s."""
    from oracle import primitives as P
    from oracle.fgp_oracle import OracleFastGP
    shift = np.random.Generator(np.random.PCG64(seed)).random(d)
    x = P.lattice_points(gen_vec(d), shift, 0, n)
    o = OracleFastGP("lattice", x, alpha=2)
    y = f_synth(torch.from_numpy(x))
    o.add_y(y)
    o.k1parts()
    o.fit(iterations=1, stop_crit_wait_iterations=10 ** 6, store_hist=False)  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    o.fit(iterations=iters, stop_crit_wait_iterations=10 ** 6, store_hist=False)
    t_fit = time.perf_counter() - t0
    out = {"fit_iters_per_s": iters / t_fit, "fit_s": t_fit, "iters": iters}
    if pm_points > 0:
        with torch.no_grad():
            c = o.coeffs().detach()
            xt = torch.rand((pm_points, d), generator=torch.Generator().manual_seed(17))
            t0 = time.perf_counter()
            o.post_mean(xt, coeffs=c)
            t_pm = time.perf_counter() - t0
        out.update(post_mean_pts_per_s=pm_points / t_pm, post_mean_s=t_pm, post_mean_points=pm_points)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--log2m", type=int, default=16, help="log2 of post_mean test points per GPU per repetition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, d, K, W = 1 << args.log2n, args.d, args.steps, max(args.warmup, 3)
    workload = "FastGPLattice d=%d n=2^%d alpha=2: fit() MLL+grad iterations and post_mean on 2^%d test points per GPU (BASELINE.json configs[2])" % (d, args.log2n, args.log2m)
    hbm_peak, peak_src = peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        iters = max(1, min(K, 8))
        torch.set_num_threads(os.cpu_count() or 1)
        r = cpu_reference_run(n, d, iters, 32)
        line = {"impl": "reference", "metric": "fit_mll_grad_iters_per_s", "value": r["fit_iters_per_s"], "unit": "iterations/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 / r["fit_iters_per_s"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "note": "reference algorithm CPU port (oracle/fgp_oracle.py: torch float64, materialised parts, log2(n)-pass transforms, autograd, Rprop); bounded sample of %d iterations" % iters},
                "cpu_baseline": {"value": r["fit_iters_per_s"], "unit": "iterations/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": "%d fit iterations at n=2^%d d=%d; post_mean on %d points" % (iters, args.log2n, d, r.get("post_mean_points", 0))},
                "e2e": {"value": r["fit_iters_per_s"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "post_mean": {"value": r.get("post_mean_pts_per_s"), "unit": "points/s"}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------- our arm
    import fastgaussianprocesses_b200 as fgp
    from fastgaussianprocesses_b200 import _lib as L
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # NCCL prints its version banner to STDOUT at every level from VERSION up (WARN included): stdout must stay one JSON line
        if os.environ.get("FGP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["FGP_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # independent GP per rank (another randomisation of the same lattice): weak scaling, no data-path collective
    seq = fgp.Lattice(d, seed=7 + rank, generating_vector=gen_vec(d))
    gp = fgp.FastGPLattice(seq, device=dev)
    x = gp.get_x_next(n)
    y_dev = f_synth(x)
    y_host = y_dev.cpu().pin_memory()
    gp.add_y_next(y_dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)  # 256 MiB > 126 MB L2
    stepper = gp.fit_stepper()
    for _ in range(W):
        stepper.step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_mark0 = time.time()
    # ---- value: K steps, each timed alone, L2 flushed between steps
    l0 = stepper.launches
    evs = []
    barrier()
    for _ in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stepper.step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_cold = max_over_ranks(sum(step_ms) * 1e-3)
    launches = stepper.launches - l0
    # ---- warm: the same K steps back to back (the natural fit() loop, working set L2-resident)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        stepper.step()
    e1.record()
    barrier()
    t_warm = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    kern = stepper.kernel_times(reps=10, flush=flush)
    stepper.close()
    # ---- e2e: public API, host buffers
    gp2 = fgp.FastGPLattice(fgp.Lattice(d, seed=7 + rank, generating_vector=gen_vec(d)), device=dev)
    gp2.get_x_next(n)
    gp2.add_y_next(y_host)
    gp2.fit(iterations=W, verbose=0, stop_crit_wait_iterations=W + 1)  # warm-up
    gp3 = fgp.FastGPLattice(fgp.Lattice(d, seed=7 + rank, generating_vector=gen_vec(d)), device=dev)
    gp3.get_x_next(n)
    barrier()
    t0 = time.perf_counter()
    gp3.add_y_next(y_host)  # H2D of this job's inputs
    data = gp3.fit(iterations=K, verbose=0, stop_crit_wait_iterations=K + 1, store_loss_hist=True)
    hyp_host = [gp3.scale.detach().cpu(), gp3.lengthscales.detach().cpu(), data["loss_hist"].cpu()]  # D2H of the result
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    fit_iters = int(data["iterations"])
    # ---- post_mean: m points per GPU, sharded test set
    m = 1 << args.log2m
    xt_host = torch.rand((m, d), generator=torch.Generator().manual_seed(17 + rank)).pin_memory()
    xt = xt_host.to(dev)
    gp.coeffs
    gp.post_mean(xt[:1024])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lp0 = L.launch_count()
    e0.record()
    pm = gp.post_mean(xt)
    if dist is not None:
        gathered = torch.empty((world, m), device=dev)
        dist.all_gather_into_tensor(gathered, pm.contiguous())
    e1.record()
    barrier()
    lp1 = L.launch_count()
    t_pm = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    barrier()
    t0 = time.perf_counter()
    pm2 = gp.post_mean(xt_host).cpu()
    torch.cuda.synchronize()
    t_pm_e2e = max_over_ranks(time.perf_counter() - t0)
    t_mark1 = time.time()
    clocks = sampler.stop(t_mark0, t_mark1) if sampler is not None else None
    # FP64 peak probe (not in MEASURED_PEAKS.json)
    fl = L.fp64_peak_probe(2000, dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fl = L.fp64_peak_probe(4000, dev)
    e1.record()
    torch.cuda.synchronize()
    fp64_peak = fl / (e0.elapsed_time(e1) * 1e-3) / 1e12
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel of the step
    alg_bytes_iter = stepper.algorithmic_bytes
    kern = [k for k in kern if k["alg_bytes"] > 0] or kern
    tmax = max(k["ms"] for k in kern)
    dom = max((k for k in kern if k["ms"] >= 0.9 * tmax), key=lambda k: k["alg_bytes"])  # ties within 10 %: the one moving more bytes
    roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9 / hbm_peak,
            # DRAM read+write bytes of that kernel per launch from the committed ncu --set full capture (same workload only)
            "traffic": NCU_TRAFFIC.get((dom["name"], args.log2n, d)), "traffic_source": "profiles/r1j_ncu_full_mll_passABC_halfspectrum_raw.csv",
            "peak_source": peak_src,
            "alg_bytes_per_launch": dom["alg_bytes"], "kernels": kern,
            "step": {"alg_bytes": alg_bytes_iter, "achieved": alg_bytes_iter / (t_cold / K) / 1e9, "frac": alg_bytes_iter / (t_cold / K) / 1e9 / hbm_peak}}
    slots = (5 * d + 1) * float(m) * n  # FP64 issue slots of the alpha=2 inner loop
    line = {"metric": "fit_mll_grad_iters_per_s", "value": world * K / t_cold, "unit": "iterations/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_cold / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2": "256 MiB buffer written between timed steps (value); e2e and warm run the natural back-to-back fit() loop",
                       "parallelism": "replicas (one independent GP per GPU); post_mean test points sharded, one NCCL all_gather" if world > 1 else "1 GPU"},
            "warm": {"value": world * K / t_warm, "unit": "iterations/s", "ms_per_step": 1e3 * t_warm / K},
            "e2e": {"value": world * fit_iters / t_e2e, "unit": "iterations/s", "h2d_bytes_per_step": int(y_host.numel() * 8 / max(fit_iters, 1)),
                    "d2h_bytes_per_step": int(sum(t.numel() for t in hyp_host) * 8 / max(fit_iters, 1)), "seconds": t_e2e, "iterations": fit_iters,
                    "note": "add_y_next(host y) + fit(iterations=K) + hyperparameters/loss history to host; copies amortised over the K iterations of the job"},
            "gpu_launches": int(launches), "roofline": roof,
            "post_mean": {"value": world * m / t_pm, "unit": "points/s", "m_per_gpu": m, "ms": 1e3 * t_pm, "launches": int(lp1 - lp0),
                          "e2e": {"value": world * m / t_pm_e2e, "unit": "points/s", "h2d_bytes": int(m * d * 8), "d2h_bytes": int(m * 8)},
                          "roofline": {"bound": "fp64", "achieved": 2 * slots / t_pm / 1e12, "peak": fp64_peak, "unit": "TFLOP/s (FP64 issue slots x2)",
                                       "frac": 2 * slots / t_pm / 1e12 / fp64_peak, "peak_source": "fgp_fp64_peak_probe DFMA chains, measured in this run"}},
            "clocks": clocks}
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        # bounded sample of the same workload, ~10-20 s of CPU work on the GPU box's host cores
        r = cpu_reference_run(n, d, 60, 128)
        line["cpu_baseline"] = {"value": r["fit_iters_per_s"], "unit": "iterations/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "60 fit iterations of the oracle port at n=2^%d d=%d (%.1f s); post_mean on 128 points (%.1f s)" % (args.log2n, d, r["fit_s"], r.get("post_mean_s", 0.0)),
                                "post_mean_points_per_s": r.get("post_mean_pts_per_s")}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
