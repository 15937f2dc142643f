"""How reproducible are the reference's own numbers?  Runs the UNMODIFIED reference (/root/reference/fastgps) twice on every
single-task fixture configuration of tests/golden/: once on the stand-in transforms used for the fixtures (radix-2 doubling
recursion, util.py:121-126) and once on mathematically identical transforms evaluated in another order (torch.fft on the
bit-reversed input; FWHT stages top-down).  The relative spread between the two is the floor below which an assert against
the fixtures says nothing about the implementation under test.  Build container only (needs /root/reference):

    python tools/reference_spread.py > profiles/r2_reference_spread.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "qmcpy_standin"))
sys.path.insert(0, "/root/reference")
torch.set_default_dtype(torch.float64)
torch.set_num_threads(4)

import fastgps  # noqa: E402
import qmcpy  # noqa: E402
from oracle import primitives as P  # noqa: E402


def bitrev_perm(n):
    m = n.bit_length() - 1
    i = np.arange(n)
    r = np.zeros(n, dtype=np.int64)
    for b in range(m):
        r |= ((i >> b) & 1) << (m - 1 - b)
    return torch.from_numpy(r)


def fftbr_alt(x):
    n = x.size(-1)
    return torch.fft.fft(x.to(torch.complex128)[..., bitrev_perm(n)], norm="ortho")


def ifftbr_alt(x):
    n = x.size(-1)
    return torch.fft.ifft(x.to(torch.complex128), norm="ortho")[..., bitrev_perm(n)]


def fwht_alt(x):
    n = x.size(-1)
    m = n.bit_length() - 1
    y = x
    batch = y.shape[:-1]
    for s in range(m - 1, -1, -1):  # top-down stage order (the stand-in runs bottom-up)
        h = 1 << s
        y = y.reshape(*batch, n // (2 * h), 2, h)
        a, b = y[..., 0, :], y[..., 1, :]
        y = torch.stack([a + b, a - b], dim=-2) / np.sqrt(2)
        y = y.reshape(*batch, n)
    return y


def evaluate(g, alt):
    fam, d, n, alpha = str(g["family"]), int(g["d"]), int(g["n"]), int(g["alpha"])
    saved = (qmcpy.fftbr_torch, qmcpy.ifftbr_torch, qmcpy.fwht_torch)
    if alt:
        qmcpy.fftbr_torch, qmcpy.ifftbr_torch, qmcpy.fwht_torch = fftbr_alt, ifftbr_alt, fwht_alt
    try:
        kw = dict(alpha=alpha, scale=torch.from_numpy(g["scale0"]), lengthscales=torch.from_numpy(g["lengthscales0"]), noise=torch.from_numpy(g["noise0"]))
        if fam == "lattice":
            seq = qmcpy.Lattice(dimension=d, generating_vector=g["z"], shift=g["shift"])
            gp = fastgps.FastGPLattice(seq, **kw)
        else:
            seq = qmcpy.DigitalNetB2(dimension=d, generating_matrices=g["C"], dshift=g["dshift"], t=int(g["t"]))
            gp = fastgps.FastGPDigitalNetB2(seq, **kw)
        x = gp.get_x_next(n)
        assert np.array_equal(x.numpy(), g["x"])
        gp.add_y_next(torch.from_numpy(g["y"]))
        xt = torch.from_numpy(g["xtest"])
        os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
        norm_term, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
        loss = 0.5 * (norm_term.sum() + logdet.sum() + n * np.log(2 * np.pi))
        loss.backward()
        out = {"loss0": loss.item(), "norm_term0": norm_term.detach().numpy().copy(), "logdet0": logdet.detach().numpy().copy(),
               "grad_raw_scale0": gp.raw_scale.grad.numpy().copy(), "grad_raw_lengthscales0": gp.raw_lengthscales.grad.numpy().copy()}
        gp.zero_grad()
        del os.environ["FASTGP_FORCE_RECOMPILE"]
        out["lam0"] = gp.get_lam(0, 0).detach().numpy()
        out["coeffs0"] = gp.coeffs.detach().numpy()
        out["pmean0"] = gp.post_mean(xt).numpy()
        out["pvar0"] = gp.post_var(xt).numpy()
        data = gp.fit(iterations=int(g["fit_iterations"]), verbose=0, store_hists=True)
        out["loss_hist"] = data["loss_hist"].numpy()
        out["lengthscales_hist"] = data["lengthscales_hist"].numpy()
        out["pmean1"] = gp.post_mean(xt).numpy()
        out["pvar1"] = gp.post_var(xt).numpy()
        return out
    finally:
        qmcpy.fftbr_torch, qmcpy.ifftbr_torch, qmcpy.fwht_torch = saved


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return None
    den = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max() / den)


def main():
    gdir = os.path.join(ROOT, "tests", "golden")
    report = {"what": "unmodified reference, stand-in transforms (fixtures) vs the same reference on torch.fft / top-down FWHT transforms: max |a-b| / max |b|",
              "cases": {}}
    for f in sorted(os.listdir(gdir)):
        if not f.endswith(".npz") or f.startswith(("mt_", "dv_", "sg_")):
            continue
        g = dict(np.load(os.path.join(gdir, f)))
        base = evaluate(g, alt=False)
        alt = evaluate(g, alt=True)
        row = {}
        for k in base:
            row[k] = {"alt_vs_fixture_order": relerr(alt[k], base[k]), "rerun_vs_committed_fixture": relerr(base[k], g[k]) if k in g else None}
        report["cases"][f[:-4]] = row
    worst = {}
    for row in report["cases"].values():
        for k, v in row.items():
            if v["alt_vs_fixture_order"] is not None:
                worst[k] = max(worst.get(k, 0.0), v["alt_vs_fixture_order"])
    report["worst_over_cases"] = worst
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
