"""StandardGP (dense GP, standard_gp.py:11-439) against fixtures produced by the reference's own StandardGP on the qmcpy
stand-in (tests/golden/make_golden.py --standard-only).  The class is plain torch (not the CUDA hot path), so the same test
runs on the CPU here and on cuda:0 on the GPU box.

Tolerances: loss / logdet 1e-10; quantities that go through the explicit inverse of a Gram matrix with condition number
~1e7..1e9 (coeffs, gradients, posterior) are compared at 1e-6 relative on the CPU -- two LAPACK evaluation orders of the same
formula already differ at that level -- and 1e-5 on CUDA (cuSOLVER)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
import os

SG_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("sg_") and f.endswith(".npz"))


def rel(a, b):
    """max |a-b| / max |b|; NaNs must sit at the same places (the reference's Matern kernels have a NaN lengthscale gradient:
    autograd of sqrt at the zero distances on the diagonal)."""
    a, b = torch.as_tensor(a).detach().cpu(), torch.as_tensor(b).detach().cpu()
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    a, b = torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def build(g, device):
    import fastgaussianprocesses_b200 as fgp
    d, T, solo = int(g["d"]), int(g["T"]), bool(g["solo"])
    batch = tuple(int(v) for v in g["batch"])
    kw = dict(shape_batch=torch.Size(batch), shape_lengthscales=torch.Size(batch + (d,))) if batch else {}
    xs = [torch.from_numpy(g["x_%d" % l]) for l in range(T)]
    ys = [torch.from_numpy(g["y_%d" % l]) for l in range(T)]
    gp = fgp.StandardGP(d, num_tasks=None if solo else T, kernel_class=str(g["kernel_class"]), noise=float(g["noise0"]), device=device,
                        data={"x": xs if not solo else xs[0], "y": ys if not solo else ys[0]}, **kw)
    return gp


def check(case, device, tol):
    g = load_golden(case)
    gp = build(g, device)
    xt = torch.from_numpy(g["xtest"]).to(device)
    ns = [int(v) for v in g["ns"]]
    norm, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
    loss = 0.5 * (norm.sum() + logdet.sum() + sum(ns) * np.log(2 * np.pi))
    assert abs(float(loss) - float(g["loss0"])) <= 1e-9 * abs(float(g["loss0"]))
    assert rel(logdet, g["logdet0"]) < 1e-10
    loss.backward()
    assert rel(gp.raw_scale.grad, g["grad_raw_scale0"]) < tol
    assert rel(gp.raw_lengthscales.grad, g["grad_raw_lengthscales0"]) < tol
    gp.zero_grad()
    with torch.no_grad():
        assert rel(gp.coeffs, g["coeffs0"]) < tol
        pm, pv = gp.post_mean(xt), gp.post_var(xt)
        assert pm.shape == g["pmean0"].shape and pv.shape == g["pvar0"].shape
        assert rel(pm, g["pmean0"]) < tol
        assert float((pv.cpu() - torch.from_numpy(g["pvar0"])).abs().max()) < tol * float(np.abs(g["pvar0"]).max() + 1)
        pc = gp.post_cov(xt[:8], xt[:5])
        assert pc.shape == g["pcov0"].shape and float((pc.cpu() - torch.from_numpy(g["pcov0"])).abs().max()) < tol * float(np.abs(g["pcov0"]).max() + 1)
        if "pcov0_eq" in g:
            pe = gp.post_cov(xt[:6], xt[:6])
            assert pe.shape == g["pcov0_eq"].shape and float((pe.cpu() - torch.from_numpy(g["pcov0_eq"])).abs().max()) < tol * float(np.abs(g["pcov0_eq"]).max() + 1)
        if "pcmean0" in g:
            assert rel(gp.post_cubature_mean(), g["pcmean0"]) < tol
            assert float((gp.post_cubature_var().cpu() - torch.from_numpy(g["pcvar0"])).abs().max()) < tol * 10
        if "pccov0" in g:
            assert float((gp.post_cubature_cov().cpu() - torch.from_numpy(g["pccov0"])).abs().max()) < tol * 10
        numer, denom = gp.get_inv_log_det_cache().get_gcv_numer_denom()
        assert abs(float((numer / denom).sum()) - float(g["gcv_loss0"])) <= tol * abs(float(g["gcv_loss0"]))
    data = gp.fit(iterations=int(g["fit_iterations"]), verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    assert data["iterations"] == int(g["fit_last_iteration"])
    assert rel(data["loss_hist"], g["loss_hist"]) < tol
    assert rel(data["lengthscales_hist"], g["lengthscales_hist"]) < 100 * tol
    assert rel(gp.post_mean(xt), g["pmean1"]) < 1000 * tol


@pytest.mark.parametrize("case", SG_CASES)
def test_standard_gp_matches_reference_fixture_cpu(case):
    check(case, "cpu", 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("case", SG_CASES)
def test_standard_gp_matches_reference_fixture_cuda(case):
    check(case, "cuda:0", 1e-5)


def test_standard_gp_api_surface():
    """Constructor defaults and the generator-driven work flow of the reference doctest (standard_gp.py:28-49): points from
    the default design, interpolation at the data, future-n variance equal to the variance after adding the points."""
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.StandardGP(2, device="cpu", seed_for_seq=7)
    assert gp.kernel_class == "gaussian" and gp.adaptive_nugget and float(gp.noise) == pytest.approx(1e-4)
    f = lambda x: torch.cos(2 * np.pi * x).sum(1)
    x = gp.get_x_next(24)
    assert x.shape == (24, 2) and float(x.min()) >= 0 and float(x.max()) < 1
    gp.add_y_next(f(x))
    assert rel(gp.post_mean(gp.x), gp.y) < 0.3  # unfitted default hyperparameters (the reference doctest reports 5e-2 at n=64)
    xt = torch.rand((16, 2), generator=torch.Generator().manual_seed(3))
    pv_future = gp.post_var(xt, n=40)
    pcv_future = gp.post_cubature_var(n=40)
    x2 = gp.get_x_next(40)
    assert x2.shape == (16, 2)
    gp.add_y_next(f(x2))
    assert torch.allclose(gp.post_var(xt), pv_future) and torch.allclose(gp.post_cubature_var(), pcv_future)
    pm, pvar, q, lo, hi = gp.post_ci(xt)
    assert (hi >= lo).all()
    assert torch.allclose(gp.post_cov(xt, xt).diagonal(), gp.post_var(xt), atol=1e-9)
    data = gp.fit(iterations=5, verbose=0)
    assert list(data.keys()) == ["iterations"]
    with pytest.raises(AssertionError):
        fgp.StandardGP(2, device="cpu", kernel_class="rbf")
